"""The reference's finest seam: `attn_implementation=<name>` (run_pseudo_labelling.py:64,230) selects
`ALL_ATTENTION_FUNCTIONS[name]`, which WhisperAttention.forward calls at HF/models/whisper/modeling_whisper.py:342-352.

    from transformers import AttentionInterface
    from kotoba_whisper_b200.attention_plugin import kwb200_attention_forward
    AttentionInterface.register("kwb200", kwb200_attention_forward)
    model = WhisperForConditionalGeneration.from_pretrained(..., attn_implementation="kwb200")

lets HF's own layer loop run on `kw_attention` (tcgen05 flash kernel for bf16, exact SIMT kernel for fp32).  Only the
mask-free case is served (encoder self-attention and decoder cross-attention); anything with a mask or dropout is handed
to HF's sdpa implementation, exactly like a partial plug-in would."""
from __future__ import annotations

import torch

from . import _lib


def kwb200_attention_forward(module, query, key, value, attention_mask=None, dropout: float = 0.0, scaling=None,
                             **kwargs):
    """query [B, H, Tq, 64], key / value [B, H, Tk, 64] -> (out [B, Tq, H, 64], None)."""
    B, H, Tq, D = query.shape
    usable = (attention_mask is None and not dropout and D == 64 and query.is_cuda and
              query.dtype in (torch.float32, torch.bfloat16) and not kwargs.get("is_causal", False))
    if not usable:
        from transformers.integrations.sdpa_attention import sdpa_attention_forward
        return sdpa_attention_forward(module, query, key, value, attention_mask, dropout=dropout, scaling=scaling,
                                      **kwargs)
    if scaling is not None and scaling != 1.0:
        query = query * scaling
    # kernel layout: element (b, t, h, e) at b*stride_b + t*stride_t + h*64 + e
    q = query.transpose(1, 2).contiguous()
    k = key.transpose(1, 2).contiguous()
    v = value.transpose(1, 2).contiguous()
    Tk = k.shape[1]
    out = torch.empty((B, Tq, H, D), dtype=query.dtype, device=query.device)
    lib = _lib.load()
    with torch.cuda.device(query.device):
        st = torch.cuda.current_stream(query.device).cuda_stream
        _lib.check(lib.kw_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, Tq, Tk,
                                    Tq * H * D, H * D, Tk * H * D, H * D, Tq * H * D, H * D,
                                    _lib.KW_BF16 if query.dtype == torch.bfloat16 else _lib.KW_F32, st), "kw_attention")
    return out, None
