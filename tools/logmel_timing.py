"""Per-phase share of the log-mel kernel's time (needs a -DKW_LOGMEL_TIMING build of the library)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kotoba_whisper_b200 import WhisperFeatureExtractorB200, _lib
lib = _lib.load()
fe = WhisperFeatureExtractorB200(feature_size=128, device="cuda:0")
audio = torch.randn(64, 480000, device="cuda") * 0.1
for _ in range(2): fe.logmel_device(audio)
torch.cuda.synchronize(); lib.kw_mel_filterbank(-1, None)
for _ in range(3): fe.logmel_device(audio)
torch.cuda.synchronize(); lib.kw_mel_filterbank(-1, None)
