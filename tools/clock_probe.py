import os, sys, time, torch
sys.path.insert(0, "/root/repo")
from bench import KOTOBA, synth_audio
from kotoba_whisper_b200 import WhisperB200Config, WhisperB200ForConditionalGeneration, WhisperFeatureExtractorB200
from kotoba_whisper_b200.random_init import random_state_dict
B = 64; dev = torch.device("cuda", 0)
cfg = WhisperB200Config(**KOTOBA)
model = WhisperB200ForConditionalGeneration.from_state_dict(random_state_dict(cfg, 0, dev), cfg, dtype=torch.bfloat16, max_batch=B, device=dev)
fe = WhisperFeatureExtractorB200(feature_size=128, device=dev)
audio = torch.from_numpy(synth_audio(B, 1000)).to(dev)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for gap_ms in (0, 0, 20, 50, 100, 0):
    for rep in range(3):
        feats = fe.logmel_device(audio)
        model.encode(feats, return_hidden=False)
        if gap_ms:
            torch.cuda.synchronize(); time.sleep(gap_ms / 1e3)
        a = ev(); model._greedy_pass(B, [50258, 50266, 50360, 50364], 128, False); b = ev(); torch.cuda.synchronize()
    print(f"gap {gap_ms:3d} ms after encode: greedy pass {a.elapsed_time(b):.2f} ms", flush=True)
# and max_length 64 vs 128 split to see whether the first half is slower
for ml in (64, 128):
    feats = fe.logmel_device(audio); model.encode(feats, return_hidden=False)
    a = ev(); model._greedy_pass(B, [50258, 50266, 50360, 50364], ml, False); b = ev(); torch.cuda.synchronize()
    print(f"in-situ max_length {ml}: {a.elapsed_time(b):.2f} ms")
