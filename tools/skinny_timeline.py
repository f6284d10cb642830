"""Timeline of one decode-time GEMM launch (CTA 0, %globaltimer): where do the ~17 us go?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kotoba_whisper_b200 import _lib
lib = _lib.load()
F32, BF16 = _lib.KW_F32, _lib.KW_BF16
st = lambda: torch.cuda.current_stream().cuda_stream
names = {0: "start", 1: "W requested", 2: "dep wait done", 3: "stage0 landed", 4: "last MMA commit", 8: "acc visible", 9: "staged", 10: "cluster barrier", 5: "epilogue done", 7: "all done"}
for (N, K, epi, tag) in [(1280, 1280, 0, "d x d store"), (1280, 1280, 2, "d x d resid"), (1280, 5120, 2, "fc2 resid"), (3840, 1280, 0, "qkv"), (5120, 1280, 1, "fc1 gelu bf16")]:
    A = torch.randn(64, K, device="cuda").bfloat16(); W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    bias = torch.randn(N, device="cuda"); od = BF16 if epi == 1 else F32
    out = torch.zeros(64, N, device="cuda", dtype=torch.bfloat16 if epi == 1 else torch.float32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stamps = torch.zeros(64, dtype=torch.int64, device="cuda")
    lib.kw_debug_gemm_stamps(stamps.data_ptr())
    res = []
    for rep in range(4):
        if os.environ.get('KW_WARM') != '1': flush.zero_()  # evict W from L2 (decode steps stream ~1.2 GB between uses of a weight)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.kw_linear(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), 64, N, K, epi, BF16, BF16, od, 2, st()))
        e1.record(); torch.cuda.synchronize()
        s = stamps.cpu().tolist()
        res.append((e0.elapsed_time(e1) * 1e3, {names[i]: (s[i] - s[0]) / 1e3 for i in names if s[i]}))
    lib.kw_debug_gemm_stamps(None)
    print(tag, f"event time {res[-1][0]:.1f} us ->", {k: round(v, 2) for k, v in res[-1][1].items()})
    print("   k-block full-barrier times (us):", [round((s[16 + i] - s[0]) / 1e3, 2) for i in range(40) if s[16 + i]])
