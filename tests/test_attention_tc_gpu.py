"""tcgen05 flash attention (kw_attention, bf16, impl auto) vs torch fp32 softmax(q k^T) v on the same bf16 inputs.
Tolerance 1.5e-2 of the output scale (P and the output are rounded to bf16; accumulation is fp32)."""
import pytest
import torch

from kotoba_whisper_b200 import _lib

pytestmark = pytest.mark.gpu
BF16 = _lib.KW_BF16


def _attn(q, k, v, B, H, Tq, Tk, impl=0):
    lib = _lib.load()
    d = H * 64
    out = torch.zeros(B, Tq, d, dtype=torch.bfloat16, device="cuda")
    lib.kw_set_gemm_impl(impl)
    _lib.check(lib.kw_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, Tq, Tk,
                                q.stride(0), q.stride(1), k.stride(0), k.stride(1), out.stride(0), out.stride(1), BF16,
                                torch.cuda.current_stream().cuda_stream), "kw_attention")
    torch.cuda.synchronize()
    lib.kw_set_gemm_impl(0)
    return out


def _ref(q, k, v, B, H, Tq, Tk):
    qh = q.float().reshape(B, Tq, H, 64).transpose(1, 2)
    kh = k.float().reshape(B, Tk, H, 64).transpose(1, 2)
    vh = v.float().reshape(B, Tk, H, 64).transpose(1, 2)
    o = torch.softmax(qh @ kh.transpose(-1, -2), -1) @ vh
    return o.transpose(1, 2).reshape(B, Tq, H * 64)


def test_attention_tc_v_path_only():
    """q = 0 -> uniform softmax -> out = mean over keys of V: isolates the P.V MMA (MN-major V descriptor)."""
    torch.manual_seed(0)
    B, H, Tq, Tk = 1, 1, 128, 128
    q = torch.zeros(B, Tq, 64, dtype=torch.bfloat16, device="cuda")
    k = torch.randn(B, Tk, 64, device="cuda").bfloat16()
    v = torch.randn(B, Tk, 64, device="cuda").bfloat16()
    out = _attn(q, k, v, B, H, Tq, Tk)
    ref = v.float().mean(1, keepdim=True).expand(B, Tq, 64)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2, f"V path: max err {err}, out[0,0,:4]={out[0,0,:4].tolist()} ref={ref[0,0,:4].tolist()}"


def test_attention_tc_p_path_only():
    """V = one-hot of (key mod 64) -> out[r, e] = P[r, e] + P[r, e+64]: isolates S = Q K^T, the softmax and P's layout."""
    torch.manual_seed(1)
    B, H, Tq, Tk = 1, 1, 128, 128
    q = (torch.randn(B, Tq, 64, device="cuda") * 0.5).bfloat16()
    k = torch.randn(B, Tk, 64, device="cuda").bfloat16()
    v = torch.zeros(B, Tk, 64, device="cuda")
    v[0, torch.arange(Tk), torch.arange(Tk) % 64] = 1.0
    v = v.bfloat16()
    out = _attn(q, k, v, B, H, Tq, Tk)
    ref = _ref(q, k, v, B, H, Tq, Tk)
    err = (out.float() - ref).abs().max().item()
    assert err <= 1e-2, f"P path: max err {err}"


@pytest.mark.parametrize("B,H,Tq,Tk", [(1, 1, 128, 128), (1, 2, 128, 384), (2, 3, 100, 37), (1, 2, 300, 1500),
                                       (1, 20, 1500, 1500), (3, 20, 1500, 1500),
                                       # key-tile edges of the 3-slot K/V ring: exactly one tile, one key into the second
                                       # tile, 4 tiles (first slot reuse), a single query row
                                       (1, 1, 64, 64), (2, 2, 130, 65), (1, 3, 257, 200), (2, 1, 1, 449)])
def test_attention_tc_matches_torch(B, H, Tq, Tk):
    torch.manual_seed(B * 1000 + Tq + Tk)
    d = H * 64
    qkv = torch.randn(B, max(Tq, Tk), 3 * d, device="cuda")
    qkv[..., :d] *= 0.5
    qkv = qkv.bfloat16()
    q, k, v = qkv[:, :Tq, :d], qkv[:, :Tk, d:2 * d], qkv[:, :Tk, 2 * d:]   # strided views like the fused QKV buffer
    out = _attn(q, k, v, B, H, Tq, Tk)
    ref = _ref(q, k, v, B, H, Tq, Tk)
    scale = max(1.0, ref.abs().max().item())
    err = (out.float() - ref).abs().max().item()
    assert err <= 1.5e-2 * scale, f"max err {err} (scale {scale})"
    simt = _attn(q, k, v, B, H, Tq, Tk, impl=1)
    assert (out.float() - simt.float()).abs().max().item() <= 2.5e-2 * scale


def test_attention_tc_running_max_moves_every_tile():
    """Scores that keep growing along the key axis: the lazy running maximum (moved only when a tile exceeds it by 2^8)
    and the O rescale in tensor memory are exercised on every tile, not just the first."""
    torch.manual_seed(11)
    B, H, Tq, Tk = 2, 2, 256, 640
    d = H * 64
    q = (torch.randn(B, Tq, d, device="cuda") * 0.3)
    k = torch.randn(B, Tk, d, device="cuda") * 0.3
    ramp = torch.linspace(0.0, 6.0, Tk, device="cuda")[None, :, None]      # |k| grows -> later tiles dominate
    k = k + ramp * torch.sign(q.mean(dim=1, keepdim=True))
    v = torch.randn(B, Tk, d, device="cuda")
    q, k, v = q.bfloat16(), k.bfloat16(), v.bfloat16()
    out = _attn(q, k, v, B, H, Tq, Tk)
    ref = _ref(q, k, v, B, H, Tq, Tk)
    scale = max(1.0, ref.abs().max().item())
    assert (out.float() - ref).abs().max().item() <= 1.5e-2 * scale
