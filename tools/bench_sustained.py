"""Sustained (power-capped) throughput of the encoder GEMM mix: loops the four per-layer shapes for ~3 s per variant and
reports TFLOP/s with the SM clock / power seen during the loop.  Variants: 1-CTA vs 2-CTA kernel."""
import os, subprocess, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kotoba_whisper_b200 import _lib
lib = _lib.load(); F32, BF16 = _lib.KW_F32, _lib.KW_BF16
st = lambda: torch.cuda.current_stream().cuda_stream
M = 96000
shapes = [(3840, 1280, 0, BF16), (1280, 1280, 2, F32), (5120, 1280, 1, BF16), (1280, 5120, 2, F32)]
bufs = []
for (N, K, epi, od) in shapes:
    A = torch.randn(M, K, device="cuda").bfloat16(); W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = torch.randn(N, device="cuda"); out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if od == BF16 else torch.float32)
    bufs.append((A, W, b, out, N, K, epi, od))
flops = sum(2.0 * M * N * K for (N, K, _, _) in shapes)
def layer():
    for (A, W, b, out, N, K, epi, od) in bufs:
        _lib.check(lib.kw_linear(A.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, epi, BF16, BF16, od, 2, st()))
def sample(stop, acc):
    while not stop.is_set():
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
        try:
            c, p = [float(x) for x in r.split(",")]; acc.append((c, p))
        except Exception:
            pass
        time.sleep(0.2)
for name, two in (("1-CTA", 0), ("2-CTA", 1), ("1-CTA", 0), ("2-CTA", 1)):
    lib.kw_set_gemm_2cta(two)
    for _ in range(20): layer()
    torch.cuda.synchronize()
    stop, acc = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, acc)); th.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); n = 900
    for _ in range(n): layer()
    b.record(); torch.cuda.synchronize(); stop.set(); th.join()
    ms = a.elapsed_time(b)
    cl = sorted(x[0] for x in acc); pw = sorted(x[1] for x in acc)
    print(f"{name}: {flops * n / ms / 1e9:7.1f} TFLOP/s sustained over {ms/1e3:.1f} s; sm clock median {cl[len(cl)//2] if cl else 0:.0f} MHz, power median {pw[len(pw)//2] if pw else 0:.0f} W", flush=True)
