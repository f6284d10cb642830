"""Helpers for the -m gpu tests: build the product model and the oracle from the same seeded HF state_dict."""
from __future__ import annotations

import functools

import torch

from _hf import build_hf
from oracle.whisper_ref import ArchConfig, WhisperRef


@functools.lru_cache(maxsize=2)
def state_dict_for(arch_items: tuple, seed: int = 0):
    """Random-init weights of the named architecture (torch.manual_seed(seed), HF initialisation), CPU fp32."""
    hf = build_hf(dict(arch_items), seed=seed)
    sd = {k: v.detach().clone() for k, v in hf.state_dict().items()}
    return sd, hf.config


def build_pair(arch: dict, dtype=torch.float32, max_batch: int = 8, oracle_weights: str = "same"):
    """-> (product model on cuda:0, oracle).  oracle_weights="rounded": oracle sees bf16-rounded matrices."""
    from kotoba_whisper_b200 import WhisperB200ForConditionalGeneration
    sd, cfg = state_dict_for(tuple(sorted(arch.items())))
    model = WhisperB200ForConditionalGeneration.from_state_dict(sd, cfg, dtype=dtype, max_batch=max_batch, device="cuda:0")
    if oracle_weights == "rounded":
        sd = {k: (v.to(torch.bfloat16).to(torch.float32) if v.dim() >= 2 and "embed_positions" not in k else v)
              for k, v in sd.items()}
    ref = WhisperRef(sd, ArchConfig(**arch))
    return model, ref
