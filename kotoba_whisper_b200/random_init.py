"""Random-init Whisper weights under HF parameter names, generated directly on the target device (no checkpoint, no
transformers import): N(0, 0.02) matrices / embeddings, zero biases, LayerNorm = (1, 0), sinusoidal encoder positions —
the initialisation HF applies to WhisperForConditionalGeneration(config) (HF/modeling_utils.py:2285-2330,
modeling_whisper.py:524-527).  Used by bench.py, where only shapes and value ranges matter."""
from __future__ import annotations

import math
from typing import Dict

import torch

from .modeling import WhisperB200Config


def random_state_dict(cfg: WhisperB200Config, seed: int = 0, device="cuda") -> Dict[str, torch.Tensor]:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    d, f = cfg.d_model, cfg.encoder_ffn_dim

    def w(*shape):
        return torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * 0.02

    def zeros(n):
        return torch.zeros(n, device=device)

    def ones(n):
        return torch.ones(n, device=device)

    sd: Dict[str, torch.Tensor] = {}
    sd["model.encoder.conv1.weight"], sd["model.encoder.conv1.bias"] = w(d, cfg.num_mel_bins, 3), zeros(d)
    sd["model.encoder.conv2.weight"], sd["model.encoder.conv2.bias"] = w(d, d, 3), zeros(d)
    inc = math.log(10000.0) / (d // 2 - 1)
    inv = torch.exp(-inc * torch.arange(d // 2, device=device))
    t = torch.arange(cfg.max_source_positions, device=device).view(-1, 1) * inv.view(1, -1)
    sd["model.encoder.embed_positions.weight"] = torch.cat([t.sin(), t.cos()], dim=1)

    def attn(p, cross=False):
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + n + ".weight"] = w(d, d)
            if n != "k_proj":
                sd[p + n + ".bias"] = zeros(d)

    def ln(p):
        sd[p + ".weight"], sd[p + ".bias"] = ones(d), zeros(d)

    for l in range(cfg.encoder_layers):
        p = f"model.encoder.layers.{l}."
        attn(p + "self_attn.")
        ln(p + "self_attn_layer_norm")
        ln(p + "final_layer_norm")
        sd[p + "fc1.weight"], sd[p + "fc1.bias"] = w(f, d), zeros(f)
        sd[p + "fc2.weight"], sd[p + "fc2.bias"] = w(d, f), zeros(d)
    ln("model.encoder.layer_norm")
    sd["model.decoder.embed_tokens.weight"] = w(cfg.vocab_size, d)
    sd["model.decoder.embed_positions.weight"] = w(cfg.max_target_positions, d)
    for l in range(cfg.decoder_layers):
        p = f"model.decoder.layers.{l}."
        attn(p + "self_attn.")
        attn(p + "encoder_attn.")
        ln(p + "self_attn_layer_norm")
        ln(p + "encoder_attn_layer_norm")
        ln(p + "final_layer_norm")
        sd[p + "fc1.weight"], sd[p + "fc1.bias"] = w(f, d), zeros(f)
        sd[p + "fc2.weight"], sd[p + "fc2.bias"] = w(d, f), zeros(d)
    ln("model.decoder.layer_norm")
    return sd
