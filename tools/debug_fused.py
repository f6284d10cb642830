"""Bring-up helper for the fused decode kernel: one short greedy pass, fused vs per-op, tokens printed side by side."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from _gpu_util import build_pair
from _synth import KOTOBA, TINY, clips
from oracle.logmel_ref import logmel_batch_f64
from kotoba_whisper_b200 import _lib
lib = _lib.load()
arch = {"tiny": TINY, "kotoba": dict(KOTOBA, encoder_layers=1)}[sys.argv[1] if len(sys.argv) > 1 else "tiny"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ML = int(sys.argv[3]) if len(sys.argv) > 3 else 8
ts = len(sys.argv) > 4 and sys.argv[4] == "ts"
model, _ = build_pair(arch, torch.bfloat16, max_batch=max(B, 2))
base = logmel_batch_f64(clips("UGSG", 70), 128)
mel = torch.from_numpy(np.concatenate([base * (1 + 0.01 * i) for i in range((B + 3) // 4)])[:B]).cuda()
prompt = [50258, 50266, 50360] + ([] if ts else [50364])
out = {}
for impl in (0, 1):
    lib.kw_set_decode_impl(2 if impl else 0)
    model.encode(mel, return_hidden=False)
    torch.cuda.synchronize()
    import time; t0 = time.perf_counter()
    lib.kw_launch_count(1)
    out[impl] = model._greedy_pass(B, prompt, ML, ts)
    print("impl", impl, "ms", (time.perf_counter() - t0) * 1e3, "launches", lib.kw_launch_count(0), flush=True)
    if impl == 1:  # second run: steady state (tensor maps, scratch and the cooperative launch are set up)
        model.encode(mel, return_hidden=False); torch.cuda.synchronize(); t0 = time.perf_counter()
        model._greedy_pass(B, prompt, ML, ts)
        print("impl", impl, "second run ms", (time.perf_counter() - t0) * 1e3, flush=True)
same = sum(int(np.array_equal(out[0][b], out[1][b])) for b in range(B))
print("identical rows", same, "of", B)
for b in range(min(B, 4)):
    print("per-op", out[0][b][:16].tolist()); print("fused ", out[1][b][:16].tolist())
