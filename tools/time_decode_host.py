"""Host enqueue time vs device time of the greedy pass (check_every = 0: no host sync inside the pass)."""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import KOTOBA, synth_audio  # noqa: E402
from kotoba_whisper_b200 import WhisperB200Config, WhisperB200ForConditionalGeneration, WhisperFeatureExtractorB200, _lib  # noqa: E402
from kotoba_whisper_b200.random_init import random_state_dict  # noqa: E402

B = 64
dev = torch.device("cuda", 0)
cfg = WhisperB200Config(**KOTOBA)
model = WhisperB200ForConditionalGeneration.from_state_dict(random_state_dict(cfg, 0, dev), cfg, dtype=torch.bfloat16,
                                                            max_batch=B, device=dev)
fe = WhisperFeatureExtractorB200(feature_size=128, device=dev)
audio = torch.from_numpy(synth_audio(B, 1000)).to(dev)
model.encode(fe.logmel_device(audio), return_hidden=False)
prompt = [50258, 50266, 50360, 50364]
pr = (C.c_int32 * 4)(*prompt)
tokens = torch.empty((B, 128), dtype=torch.int32, device=dev)
for rep in range(5):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    t0 = time.perf_counter()
    _lib.check(model._lib.kw_greedy_pass(model._handle, B, pr, 4, 128, 0, 0, tokens.data_ptr(), model._stream()))
    t1 = time.perf_counter()
    b.record()
    torch.cuda.synchronize()
    print(f"rep {rep}: host enqueue {1e3*(t1-t0):.2f} ms, device {a.elapsed_time(b):.2f} ms")
