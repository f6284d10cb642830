"""Building-block kernels through the C ABI vs plain torch fp32 references."""
import ctypes as C

import numpy as np
import pytest
import torch

from kotoba_whisper_b200 import _lib

pytestmark = pytest.mark.gpu
F32, BF16 = _lib.KW_F32, _lib.KW_BF16
TD = {F32: torch.float32, BF16: torch.bfloat16}


def _st():
    return torch.cuda.current_stream().cuda_stream


def _linear(A, W, bias, epi, out_dtype, impl=1, out=None):
    lib = _lib.load()
    M, K = A.shape
    N = W.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=TD[out_dtype], device="cuda")
    at = BF16 if A.dtype == torch.bfloat16 else F32
    wt = BF16 if W.dtype == torch.bfloat16 else F32
    _lib.check(lib.kw_linear(A.data_ptr(), W.data_ptr(), bias.data_ptr() if bias is not None else None, out.data_ptr(),
                             M, N, K, epi, at, wt, out_dtype, impl, _st()), "kw_linear")
    return out


@pytest.mark.parametrize("M,N,K", [(1, 64, 16), (4, 51866, 128), (64, 3840, 1280), (200, 1280, 384), (3000, 384, 1280),
                                   (129, 130, 48), (1500, 1280, 5120)])
def test_linear_simt_fp32(M, N, K):
    torch.manual_seed(M + N + K)
    A, W, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") * 0.05, torch.randn(N, device="cuda")
    ref = torch.nn.functional.linear(A.double(), W.double(), b.double())
    out = _linear(A, W, b, 0, F32)
    assert (out.double() - ref).abs().max() <= 2e-5 * max(1.0, ref.abs().max().item())
    g = _linear(A, W, b, 1, F32)
    assert (g.double() - torch.nn.functional.gelu(ref)).abs().max() <= 2e-5 * max(1.0, ref.abs().max().item())
    x0 = torch.randn(M, N, device="cuda")
    r = _linear(A, W, b, 2, F32, out=x0.clone())
    assert (r.double() - (x0.double() + ref)).abs().max() <= 2e-5 * max(1.0, ref.abs().max().item())
    nb = _linear(A, W, None, 0, F32)
    assert (nb.double() - (ref - b.double())).abs().max() <= 2e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M,N,K", [(64, 1280, 1280), (257, 640, 256), (8, 51866, 128)])
def test_linear_simt_bf16_storage(M, N, K):
    torch.manual_seed(1)
    A, W = torch.randn(M, K, device="cuda"), (torch.randn(N, K, device="cuda") * 0.05)
    b = torch.randn(N, device="cuda")
    Ab, Wb = A.bfloat16(), W.bfloat16()
    ref = torch.nn.functional.linear(Ab.double(), Wb.double(), b.double())
    out = _linear(Ab, Wb, b, 0, BF16)
    assert (out.double() - ref).abs().max() <= 1e-2 * max(1.0, ref.abs().max().item())   # bf16 output rounding
    out32 = _linear(Ab, Wb, b, 0, F32)
    assert (out32.double() - ref).abs().max() <= 2e-5 * max(1.0, ref.abs().max().item())
    mix = _linear(A, Wb, b, 0, F32)                                                        # f32 activations x bf16 weights
    ref2 = torch.nn.functional.linear(A.double(), Wb.double(), b.double())
    assert (mix.double() - ref2).abs().max() <= 2e-5 * max(1.0, ref2.abs().max().item())


@pytest.mark.parametrize("rows,d", [(1, 128), (7, 192), (64, 1280), (3001, 1280)])
def test_layernorm(rows, d):
    lib = _lib.load()
    torch.manual_seed(rows)
    x = torch.randn(rows, d, device="cuda") * 3 + 1
    w, b = torch.randn(d, device="cuda"), torch.randn(d, device="cuda")
    ref = torch.nn.functional.layer_norm(x, (d,), w, b, 1e-5)
    out = torch.empty_like(x)
    _lib.check(lib.kw_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), rows, d, F32, _st()))
    assert (out - ref).abs().max() <= 1e-5
    ob = torch.empty(rows, d, dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.kw_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), ob.data_ptr(), rows, d, BF16, _st()))
    assert (ob.float() - ref).abs().max() <= 2e-2 * ref.abs().max()


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("B,H,Tq,Tk", [(1, 2, 64, 64), (2, 3, 100, 37), (1, 20, 1500, 1500), (3, 1, 1, 130)])
def test_attention_seam(dtype, B, H, Tq, Tk):
    lib = _lib.load()
    lib.kw_set_gemm_impl(1)  # SIMT attention
    torch.manual_seed(B * 100 + Tq)
    d = H * 64
    q = (torch.randn(B, Tq, d, device="cuda") * 0.5).to(TD[dtype])
    kv = (torch.randn(B, Tk, 2 * d, device="cuda")).to(TD[dtype])  # interleaved K|V rows: exercises strides
    k, v = kv[..., :d], kv[..., d:]
    out = torch.empty(B, Tq, d, dtype=TD[dtype], device="cuda")
    es = 1
    _lib.check(lib.kw_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, Tq, Tk, Tq * d, d,
                                Tk * 2 * d, 2 * d, Tq * d, d, dtype, _st()), "kw_attention")
    lib.kw_set_gemm_impl(0)
    qh = q.float().view(B, Tq, H, 64).transpose(1, 2)
    kh = k.float().reshape(B, Tk, H, 64).transpose(1, 2)
    vh = v.float().reshape(B, Tk, H, 64).transpose(1, 2)
    ref = torch.softmax(qh @ kh.transpose(-1, -2), -1) @ vh
    ref = ref.transpose(1, 2).reshape(B, Tq, d)
    tol = 2e-5 if dtype == F32 else 1.5e-2
    assert (out.float() - ref).abs().max() <= tol * max(1.0, ref.abs().max().item())
