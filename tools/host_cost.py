import os, sys, time, torch
sys.path.insert(0, "/root/repo")
from kotoba_whisper_b200 import _lib
lib = _lib.load(); F32, BF16 = _lib.KW_F32, _lib.KW_BF16
st = torch.cuda.current_stream().cuda_stream
x = torch.randn(64, 1280, device="cuda"); w = torch.randn(1280, device="cuda"); b = torch.randn(1280, device="cuda")
o = torch.empty(64, 1280, device="cuda", dtype=torch.bfloat16)
A = torch.randn(64, 1280, device="cuda").bfloat16(); W = (torch.randn(1280, 1280, device="cuda") * 0.02).bfloat16()
W3 = (torch.randn(3840, 1280, device="cuda") * 0.02).bfloat16(); o3 = torch.zeros(64, 3840, device="cuda")
out = torch.zeros(64, 1280, device="cuda")
def t(fn, n=2000):
    for _ in range(50): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6
print("layernorm           host %.2f us/launch  (incl. drain %.2f)" % t(lambda: lib.kw_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), 64, 1280, BF16, st)))
print("skinny splitK resid host %.2f us/launch  (incl. drain %.2f)" % t(lambda: lib.kw_linear(A.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr(), 64, 1280, 1280, 2, BF16, BF16, F32, 2, st)))
print("skinny qkv store    host %.2f us/launch  (incl. drain %.2f)" % t(lambda: lib.kw_linear(A.data_ptr(), W3.data_ptr(), None, o3.data_ptr(), 64, 3840, 1280, 0, BF16, BF16, F32, 2, st)))
print("python ctypes no-op  %.2f us" % t(lambda: lib.kw_launch_count(0))[0])
