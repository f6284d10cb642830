"""Greedy-pass time alone (encoder output fixed), N repetitions: mean / min / max — for A/B of decode-side changes,
whose effect (~1 ms) is below the rep-to-rep noise of tools/time_stages.py."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import KOTOBA, synth_audio  # noqa: E402
from kotoba_whisper_b200 import WhisperB200Config, WhisperB200ForConditionalGeneration, WhisperFeatureExtractorB200  # noqa: E402
from kotoba_whisper_b200.random_init import random_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda", 0)
cfg = WhisperB200Config(**KOTOBA)
model = WhisperB200ForConditionalGeneration.from_state_dict(random_state_dict(cfg, 0, dev), cfg, dtype=torch.bfloat16,
                                                            max_batch=B, device=dev)
fe = WhisperFeatureExtractorB200(feature_size=128, device=dev)
audio = torch.from_numpy(synth_audio(B, 1000)).to(dev)
model.encode(fe.logmel_device(audio), return_hidden=False)
ts = []
for rep in range(N + 2):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = model._greedy_pass(B, [50258, 50266, 50360, 50364], 128, False)
    b.record()
    torch.cuda.synchronize()
    if rep >= 2:
        ts.append(a.elapsed_time(b))
import numpy as np  # noqa: E402
ids = np.asarray(out).astype(np.int64).ravel()
chk = int((ids * (np.arange(1, ids.size + 1) % 1009)).sum())
print(f"tokens checksum {chk}  ", end="")
print(f"greedy pass B={B}: mean {sum(ts)/len(ts):.2f} ms  min {min(ts):.2f}  max {max(ts):.2f}  ({N} reps)")
