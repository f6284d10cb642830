"""Builders for the *reference* implementation (transformers' Whisper, which is what the reference repo calls at
run_pseudo_labelling.py:224-232,338).  Used by the golden generator, by oracle-vs-HF tests and by bench.py's
reference arm.  transformers lives in site-packages (not /root/reference), so this works on the GPU box too."""
from __future__ import annotations

import torch


def build_hf(arch: dict, seed: int = 0, dtype=torch.float32):
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    from transformers.models.whisper.tokenization_whisper import LANGUAGES
    from transformers.models.whisper.configuration_whisper import NON_SPEECH_TOKENS_MULTI

    cfg = WhisperConfig(decoder_start_token_id=50258, eos_token_id=50257, bos_token_id=50257, pad_token_id=50256,
                        **arch)
    torch.manual_seed(seed)
    model = WhisperForConditionalGeneration(cfg).eval().to(dtype)
    g = model.generation_config
    g.lang_to_id = {f"<|{c}|>": 50259 + i for i, c in enumerate(LANGUAGES)}
    g.task_to_id = {"translate": 50359, "transcribe": 50360}
    g.no_timestamps_token_id = 50364
    g.prev_sot_token_id = 50362
    g.is_multilingual = True
    g.max_initial_timestamp_index = 50
    g.max_length = 448
    g.eos_token_id = g.pad_token_id = g.bos_token_id = 50257
    g.decoder_start_token_id = 50258
    g.begin_suppress_tokens = [220, 50257]
    g.suppress_tokens = NON_SPEECH_TOKENS_MULTI[:82] + [50258, 50359, 50360, 50361, 50362, 50363]
    return model
