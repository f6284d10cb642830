"""Chunked long-form transcription as the HF ASR pipeline runs it for run_speed_eval.py / run_short_form_eval.py.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates HF/pipelines/automatic_speech_recognition.py:61-84 (`chunk_iter`: windows of chunk_len samples advanced by
chunk_len - stride_left - stride_right, every window padded to 30 s by the feature extractor), :428-444 (chunk_len =
round(chunk_length_s * sr), stride = chunk_length_s / 6 each side) and the token-level merge of
HF/models/whisper/tokenization_whisper.py:1153-1270 (`_find_longest_common_sequence`), driving the oracle's own log-mel
and greedy generate.  Pure-Python loops: small cases only.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from .logmel_ref import logmel_batch_f64


def chunk_bounds(n: int, chunk_len: int, stride_left: int, stride_right: int):
    step = chunk_len - stride_left - stride_right
    out = []
    for start in range(0, n, step):
        end = start + chunk_len
        is_last = end >= n
        left = 0 if start == 0 else stride_left
        right = 0 if is_last else stride_right
        length = min(end, n) - start
        if length > left:
            out.append((start, start + length, (length, left, right)))
        if is_last:
            break
    return out


def longest_common_sequence_merge(sequences: Sequence[Sequence[int]]) -> List[int]:
    left = list(sequences[0])
    total: List[int] = []
    for right in sequences[1:]:
        right = list(right)
        ll, rl = len(left), len(right)
        best, idx = 0.0, (ll, ll, 0, 0)
        for i in range(1, ll + rl):
            eps = i / 10000.0
            l0, l1 = max(0, ll - i), min(ll, ll + rl - i)
            r0, r1 = max(0, i - ll), min(rl, i)
            matches = sum(1 for a, b in zip(left[l0:l1], right[r0:r1]) if a == b)
            score = matches / i + eps
            if matches > 1 and score > best:
                best, idx = score, (l0, l1, r0, r1)
        l0, l1, r0, r1 = idx
        total.extend(left[: (l0 + l1) // 2])
        left = right[(r0 + r1) // 2:]
    total.extend(left)
    return total


def transcribe_longform_ref(ref_model, audio: np.ndarray, n_mels: int, chunk_length_s: float = 15.0, sr: int = 16000,
                            language="ja", task="transcribe", max_length=None):
    """-> (merged token ids, per-chunk token ids) using the oracle model (`WhisperRef`)."""
    chunk_len = int(round(chunk_length_s * sr))
    stride = int(round(chunk_length_s / 6 * sr))
    spans = chunk_bounds(len(audio), chunk_len, stride, stride)
    mel = torch.from_numpy(logmel_batch_f64([audio[s:e] for s, e, _ in spans], n_mels))
    with torch.no_grad():
        ids = ref_model.generate(mel, language=language, task=task, return_timestamps=False, max_length=max_length)
    eos = ref_model.gen.eos_token_id
    per_chunk = [[t for t in row if t < eos] for row in ids.tolist()]
    return longest_common_sequence_merge([c for c in per_chunk if c]), per_chunk
