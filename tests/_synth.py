"""Seeded synthetic inputs and random-init model builders shared by tests, bench and the golden generator.

Audio families follow SURVEY.md §8(d): U = the reference's own dummy clip recipe (run_speed_eval.py:14-17),
G = gaussian, S = gaussian for the first few seconds then zeros (exercises padding / the -1.5 floor).
"""
from __future__ import annotations

import numpy as np

KOTOBA = dict(vocab_size=51866, num_mel_bins=128, d_model=1280, encoder_layers=32, decoder_layers=2,
              encoder_attention_heads=20, decoder_attention_heads=20, encoder_ffn_dim=5120, decoder_ffn_dim=5120)
TEACHER = dict(KOTOBA, decoder_layers=32)
# same token-id layout (vocab 51866) so every processor rule is exercised, but small enough for CPU tests
TINY = dict(vocab_size=51866, num_mel_bins=128, d_model=128, encoder_layers=2, decoder_layers=2,
            encoder_attention_heads=2, decoder_attention_heads=2, encoder_ffn_dim=512, decoder_ffn_dim=512)
TINY80 = dict(TINY, num_mel_bins=80, d_model=192, encoder_attention_heads=3, decoder_attention_heads=3,
              encoder_ffn_dim=384, decoder_ffn_dim=384, encoder_layers=1, decoder_layers=3)


def clip(family: str, seed: int, n: int = 480000) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if family == "U":
        return ((rng.random(n) - 0.5) * 2 * 0.007).astype(np.float32)
    if family == "G":
        return (rng.standard_normal(n) * 0.1).astype(np.float32)
    if family == "S":
        keep = int(rng.integers(5 * 16000, 25 * 16000))
        return (rng.standard_normal(keep) * 0.1).astype(np.float32)
    raise ValueError(family)


def clips(spec: str, seed0: int):
    """spec like "UGSG" -> list of clips with seeds seed0, seed0+1, ..."""
    return [clip(f, seed0 + i) for i, f in enumerate(spec)]
