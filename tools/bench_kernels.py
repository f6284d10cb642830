"""Per-kernel micro-benchmarks through the C ABI (CUDA events, warm-up, operands larger than L2 for the big shapes).
    python tools/bench_kernels.py [gemm] [skinny] [attn] [logmel]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kotoba_whisper_b200 import WhisperFeatureExtractorB200, _lib  # noqa: E402

lib = _lib.load()
F32, BF16 = _lib.KW_F32, _lib.KW_BF16
what = sys.argv[1:] or ["gemm", "skinny", "attn", "logmel"]
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def linear(M, N, K, epi, out_dtype, tag):
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if out_dtype == BF16 else torch.float32)
    fn = lambda: _lib.check(lib.kw_linear(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, epi, BF16,  # noqa: E731
                                          BF16, out_dtype, 2, st()))
    ms = timeit(fn)
    tf = 2.0 * M * N * K / ms / 1e9
    gb = (N * K * 2 + M * K * 2) / ms / 1e6
    print(f"{tag:28s} M={M:6d} N={N:6d} K={K:5d}  {ms*1e3:9.1f} us  {tf:8.1f} TFLOP/s  weights+act {gb:8.1f} GB/s", flush=True)


if "2cta" in what:
    lib.kw_set_gemm_2cta(1)
    what.append("gemm")
if "gemm" in what:
    M = 96000
    linear(M, 3840, 1280, 0, BF16, "qkv (store bf16)")
    linear(M, 1280, 1280, 2, F32, "out_proj (resid f32)")
    linear(M, 5120, 1280, 1, BF16, "fc1 (gelu bf16)")
    linear(M, 1280, 5120, 2, F32, "fc2 (resid f32)")
    linear(M, 2560, 1280, 0, BF16, "cross kv (store bf16)")
    linear(2 * M, 1280, 384, 1, BF16, "conv1 (gelu bf16)")
if "skinny" in what:
    for (N, K, epi, od, tag) in [(3840, 1280, 0, F32, "dec qkv"), (1280, 1280, 2, F32, "dec out (resid)"),
                                 (1280, 1280, 0, F32, "dec q_x"), (5120, 1280, 1, BF16, "dec fc1 (gelu)"),
                                 (1280, 5120, 2, F32, "dec fc2 (resid)"), (51866, 1280, 0, F32, "vocab")]:
        linear(64, N, K, epi, od, tag)
if "attn" in what:
    B, H, T = 64, 20, 1500
    d = H * 64
    qkv = (torch.randn(B, T, 3 * d, device="cuda") * 0.5).bfloat16()
    out = torch.zeros(B, T, d, device="cuda", dtype=torch.bfloat16)
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    fn = lambda: _lib.check(lib.kw_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, T, T,  # noqa: E731
                                             T * 3 * d, 3 * d, T * 3 * d, 3 * d, T * d, d, BF16, st()))
    ms = timeit(fn, iters=10)
    print(f"{'encoder attention':28s} B={B} H={H} T={T}  {ms*1e3:9.1f} us  {4.0*B*H*T*T*64/ms/1e9:8.1f} TFLOP/s", flush=True)
if "logmel" in what:
    for nm in (80, 128):
        fe = WhisperFeatureExtractorB200(feature_size=nm, device="cuda:0")
        for B in (64, 1024):
            audio = torch.randn(B, 480000, device="cuda") * 0.1
            ms = timeit(lambda: fe.logmel_device(audio), iters=5, warm=2)
            byts = B * (480000 * 4 + nm * 3000 * 4)
            print(f"{'log-mel':28s} n_mels={nm} B={B:5d}  {ms*1e3:9.1f} us  {byts/ms/1e6:8.1f} GB/s  {B/ms*1e3:9.0f} clips/s",
                  flush=True)
if "ln" in what:
    rows, d = 96000, 1280
    x = torch.randn(rows, d, device="cuda"); w = torch.randn(d, device="cuda"); b = torch.randn(d, device="cuda")
    out = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    fn = lambda: _lib.check(lib.kw_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), rows, d, BF16, st()))  # noqa: E731
    ms = timeit(fn, iters=20)
    print(f"{'encoder LayerNorm':28s} rows={rows} d={d}  {ms*1e3:9.1f} us  {rows*d*6/ms/1e6:8.1f} GB/s", flush=True)
