"""tcgen05 / TMEM / TMA GEMM (kw_linear impl=2) vs torch fp32 on the same bf16 operands, incl. M / N tails and every
epilogue.  fp32 accumulation: f32 outputs agree to ~1e-5 relative, bf16 outputs to bf16 rounding."""
import pytest
import torch

from kotoba_whisper_b200 import _lib

pytestmark = pytest.mark.gpu
F32, BF16 = _lib.KW_F32, _lib.KW_BF16


def _run(A, W, bias, epi, out_dtype, out=None, impl=2):
    lib = _lib.load()
    M, K = A.shape
    N = W.shape[0]
    if out is None:
        out = torch.zeros((M, N), dtype=torch.bfloat16 if out_dtype == BF16 else torch.float32, device="cuda")
    _lib.check(lib.kw_linear(A.data_ptr(), W.data_ptr(), bias.data_ptr() if bias is not None else None, out.data_ptr(),
                             M, N, K, epi, BF16, BF16, out_dtype, impl, torch.cuda.current_stream().cuda_stream), "kw_linear")
    torch.cuda.synchronize()
    return out


SHAPES = [(128, 256, 64), (256, 512, 128), (100, 128, 128), (1500, 1280, 1280), (6000, 3840, 1280), (6000, 1280, 384),
          (3000, 1280, 3840), (1500, 5120, 1280), (1500, 1280, 5120), (333, 192, 576), (4 * 1500, 2560, 1280),
          # decode-time (skinny, transposed-product kernel): M = batch <= 64, any N
          (64, 3840, 1280), (64, 1280, 5120), (4, 51866, 1280), (1, 1280, 1280), (33, 5120, 1280), (64, 128, 128),
          (17, 200, 64),
          # cluster split-K (N <= ~1500: tiles x split <= 148 CTAs): uneven k-ranges, 2- and 3-way, 8 / 4-byte store paths
          (64, 1280, 1280), (5, 1000, 832), (7, 1282, 512), (3, 333, 1280), (40, 1536, 1024),
          # 40-row tiles (32-row tiles would need a second wave on 148 SMs): exact and ragged last tile, 8 / 4-byte stores
          (64, 5120, 1280), (7, 5040, 256), (5, 4762, 128),
          # 65..128 batch rows (coalesced decode batches): the same kernel with the batch as a 128-wide MMA N; every tile
          # height (32 with 2- / 3-way split-K, 40, 128), ragged batch and feature tails
          (128, 1280, 1280), (128, 3840, 1280), (128, 5120, 1280), (128, 1280, 5120), (128, 51866, 1280), (65, 1280, 1280),
          (100, 1282, 512), (97, 5040, 256), (127, 333, 1280), (96, 51866, 1280)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_tc_matches_torch(M, N, K):
    torch.manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, device="cuda")).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda")
    ref = torch.nn.functional.linear(A.float(), W.float(), b)
    scale = max(1.0, ref.abs().max().item())
    o32 = _run(A, W, b, 0, F32)
    assert (o32 - ref).abs().max().item() <= 3e-5 * scale, "store f32"
    o16 = _run(A, W, b, 0, BF16)
    assert (o16.float() - ref).abs().max().item() <= 1e-2 * scale, "store bf16"
    g16 = _run(A, W, b, 1, BF16)
    assert (g16.float() - torch.nn.functional.gelu(ref)).abs().max().item() <= 1e-2 * scale, "gelu bf16"
    x0 = torch.randn(M, N, device="cuda")
    r32 = _run(A, W, b, 2, F32, out=x0.clone())
    assert (r32 - (x0 + ref)).abs().max().item() <= 3e-5 * scale, "residual f32"
    nb = _run(A, W, None, 0, F32)
    assert (nb - (ref - b)).abs().max().item() <= 3e-5 * scale, "no bias"
    # same answer as the SIMT kernel (the exact-fp32-accumulation reference path)
    s32 = _run(A, W, b, 0, F32, impl=1)
    assert (o32 - s32).abs().max().item() <= 3e-5 * scale


def test_skinny_split_k_is_deterministic():
    """The split-K partials are added in rank order by one CTA (no atomics): repeated launches give identical bits."""
    torch.manual_seed(5)
    for M in (64, 128):
        A = torch.randn(M, 5120, device="cuda").bfloat16()
        W = (torch.randn(1280, 5120, device="cuda") * 0.05).bfloat16()
        b = torch.randn(1280, device="cuda")
        x0 = torch.randn(M, 1280, device="cuda")
        first = _run(A, W, b, 2, F32, out=x0.clone())
        for _ in range(5):
            assert torch.equal(_run(A, W, b, 2, F32, out=x0.clone()), first)


def test_skinny_rows_do_not_depend_on_batch_width():
    """A batch row's result is the same bits whether it is computed in a 64-wide or a 128-wide launch (same k order, fp32
    accumulation per output element): coalescing two decode batches into one cannot change a token."""
    torch.manual_seed(11)
    for N, K, epi, ot in ((1280, 1280, 2, F32), (3840, 1280, 0, F32), (5120, 1280, 1, BF16), (1280, 5120, 2, F32), (51866, 1280, 0, F32)):
        A = torch.randn(128, K, device="cuda").bfloat16()
        W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        b = torch.randn(N, device="cuda")
        x0 = torch.randn(128, N, device="cuda") if epi == 2 else None
        wide = _run(A, W, b, epi, ot, out=x0.clone() if x0 is not None else None)
        lo = _run(A[:64].contiguous(), W, b, epi, ot, out=x0[:64].clone() if x0 is not None else None)
        hi = _run(A[64:].contiguous(), W, b, epi, ot, out=x0[64:].clone() if x0 is not None else None)
        assert torch.equal(wide[:64], lo) and torch.equal(wide[64:], hi), (N, K, epi)


def test_gemm_tc_full_batch_shape_linearity():
    """configs[1] size (M = 64 x 1500): property check instead of a dense reference — linearity in A and row independence."""
    torch.manual_seed(0)
    M, N, K = 96000, 1280, 1280
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    full = _run(A, W, None, 0, F32)
    idx = torch.tensor([0, 127, 128, 5000, 77777, 95999], device="cuda")
    ref = torch.nn.functional.linear(A[idx].float(), W.float())
    assert (full[idx] - ref).abs().max().item() <= 3e-5 * ref.abs().max().item()
    twice = _run((A.float() * 2).bfloat16(), W, None, 0, F32)   # x2 is exact in bf16
    assert torch.equal(twice, full * 2)


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 512, 256), (1500, 1280, 1280), (6000, 3840, 1280), (6000, 1280, 384),
                                   (3000, 1280, 3840), (1500, 5120, 1280), (333, 192, 576), (4 * 1500, 2560, 1280)])
def test_gemm_tc_2cta_matches_torch(M, N, K):
    """cta_group::2 kernel (CTA pairs, 256 x 256 tiles, M = 256 MMAs): same checks as the 1-CTA kernel."""
    lib = _lib.load()
    lib.kw_set_gemm_2cta(1)
    try:
        torch.manual_seed(M + N + K)
        A = torch.randn(M, K, device="cuda").bfloat16()
        W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        b = torch.randn(N, device="cuda")
        ref = torch.nn.functional.linear(A.float(), W.float(), b)
        scale = max(1.0, ref.abs().max().item())
        o32 = _run(A, W, b, 0, F32)
        assert (o32 - ref).abs().max().item() <= 3e-5 * scale, "store f32"
        g16 = _run(A, W, b, 1, BF16)
        assert (g16.float() - torch.nn.functional.gelu(ref)).abs().max().item() <= 1e-2 * scale, "gelu bf16"
        x0 = torch.randn(M, N, device="cuda")
        r32 = _run(A, W, b, 2, F32, out=x0.clone())
        assert (r32 - (x0 + ref)).abs().max().item() <= 3e-5 * scale, "residual f32"
    finally:
        lib.kw_set_gemm_2cta(0)
