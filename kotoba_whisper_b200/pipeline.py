"""Chunked long-form transcription as `run_speed_eval.py` / `run_short_form_eval.py` drive it through the HF ASR
pipeline: 15 s windows with a 2.5 s stride on each side, every window padded to 30 s, one greedy
`generate` per batch of windows (no timestamps), then a token-level longest-common-sequence merge of neighbouring windows.

Restates HF/pipelines/automatic_speech_recognition.py:61-84 (`chunk_iter`), :428-444 (chunk / stride sizes) and
HF/models/whisper/tokenization_whisper.py:1153-1270 (`_find_longest_common_sequence`).  The merge works on token ids, so
no tokenizer is needed; special ids (>= eos) are dropped before merging, as `_decode_asr` does for text-only output.
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Sequence

import numpy as np
import torch


def chunk_spans(n_samples: int, chunk_len: int, stride_left: int, stride_right: int):
    """Yield (start, end, (chunk_len, left, right), is_last) exactly as the pipeline's chunk_iter walks the audio."""
    step = chunk_len - stride_left - stride_right
    if step <= 0:
        raise ValueError("chunk_len must exceed stride_left + stride_right")
    for start in range(0, n_samples, step):
        end = min(start + chunk_len, n_samples)
        is_last = start + chunk_len >= n_samples
        left = 0 if start == 0 else stride_left
        right = 0 if is_last else stride_right
        if end - start > left:
            yield start, end, (end - start, left, right), is_last
        if is_last:
            break


def chunk_iter(inputs: np.ndarray, feature_extractor, chunk_len: int, stride_left: int, stride_right: int,
               dtype=None) -> Iterator[dict]:
    """Same contract as the HF pipeline helper: one dict per window with `input_features`, `attention_mask`, `stride`."""
    for start, end, stride, is_last in chunk_spans(inputs.shape[0], chunk_len, stride_left, stride_right):
        processed = feature_extractor(inputs[start:end], sampling_rate=feature_extractor.sampling_rate,
                                      return_tensors="pt", return_attention_mask=True)
        if dtype is not None:
            processed = processed.to(dtype=dtype)
        yield {"is_last": is_last, "stride": stride, **processed}


def _merge_pair_indices(left: Sequence[int], right: Sequence[int]):
    """Best sliding overlap between the tail of `left` and the head of `right` (fraction of equal tokens, > 1 match,
    longer overlaps favoured by i/10000) -> (left_start, left_stop, right_start, right_stop)."""
    ll, rl = len(left), len(right)
    la, ra = np.asarray(left, dtype=np.int64), np.asarray(right, dtype=np.int64)
    best, best_idx = 0.0, (ll, ll, 0, 0)
    for i in range(1, ll + rl):
        l0, l1 = max(0, ll - i), min(ll, ll + rl - i)
        r0, r1 = max(0, i - ll), min(rl, i)
        matches = int(np.sum(la[l0:l1] == ra[r0:r1]))
        score = matches / i + i / 10000.0
        if matches > 1 and score > best:
            best, best_idx = score, (l0, l1, r0, r1)
    return best_idx


def merge_chunk_tokens(sequences: Sequence[Sequence[int]]) -> List[int]:
    """Token-level merge of consecutive windows: keep the left window up to the middle of the matched overlap and the
    right window from the middle on (left half trusted to the left window, right half to the right one)."""
    if not sequences:
        return []
    left = list(sequences[0])
    total: List[int] = []
    for right in sequences[1:]:
        right = list(right)
        l0, l1, r0, r1 = _merge_pair_indices(left, right)
        total.extend(left[: (l0 + l1) // 2])
        left = right[(r0 + r1) // 2:]
    total.extend(left)
    return total


def transcribe_longform(model, feature_extractor, audio: np.ndarray, chunk_length_s: float = 15.0,
                        stride_length_s: Optional[float] = None, batch_size: int = 64, language: Optional[str] = None,
                        task: Optional[str] = None, max_new_tokens: Optional[int] = None,
                        return_chunk_tokens: bool = False, device_chunker: bool = True, stats: Optional[dict] = None):
    """audio: mono float waveform at feature_extractor.sampling_rate -> merged text token ids (list of int).

    Window sizes follow the pipeline: chunk_len = round(chunk_length_s * sr), stride = chunk_length_s / 6 on each side.
    Device-side chunker (default): the recording is uploaded ONCE and `kw_logmel_windows` frames every 15 s window in
    place from an (offset, length) table, synthesising the zero pad to 30 s on the device — no per-window host slicing,
    staging or re-upload (the host path re-copied every window with its 15 s of zeros: ~3x the PCIe bytes).
    `device_chunker=False` keeps the per-window host path (what HF's chunk_iter does) for A/B parity.
    Windows are decoded greedily without timestamps, `batch_size` per generate; `num_beams` is 1 as in the transformers
    version the reference was written against."""
    sr = feature_extractor.sampling_rate
    audio = np.ascontiguousarray(np.asarray(audio, dtype=np.float32).reshape(-1))
    chunk_len = int(round(chunk_length_s * sr))
    stride_s = chunk_length_s / 6 if stride_length_s is None else stride_length_s
    stride = int(round(stride_s * sr))
    spans = list(chunk_spans(audio.shape[0], chunk_len, stride, stride))
    eos = model.generation_config.eos_token_id
    per_chunk: List[List[int]] = []
    h2d = 0
    if device_chunker:
        rec = torch.from_numpy(audio).to(model.device, non_blocking=True)
        h2d += audio.nbytes
    for b0 in range(0, len(spans), batch_size):
        batch = spans[b0:b0 + batch_size]
        if device_chunker:
            starts = torch.tensor([s for s, _, _, _ in batch], dtype=torch.int64)
            lens = torch.tensor([e - s for s, e, _, _ in batch], dtype=torch.int32)
            feats = feature_extractor.logmel_windows(rec, starts, lens)
            h2d += starts.numel() * 12
        else:
            feats = feature_extractor([audio[s:e] for s, e, _, _ in batch], sampling_rate=sr, return_tensors="pt",
                                      keep_on_device=True)["input_features"]
            h2d += len(batch) * feature_extractor.n_samples * 4
        kw = {} if max_new_tokens is None else {"max_new_tokens": max_new_tokens}
        ids = model.generate(feats, language=language, task=task, return_timestamps=False, **kw)
        ids = ids.cpu().tolist()
        for row in ids:
            per_chunk.append([t for t in row if t < eos])
    merged = merge_chunk_tokens([c for c in per_chunk if c])
    if stats is not None:
        stats.update(h2d_bytes=h2d, windows=len(spans))
    if return_chunk_tokens:
        return merged, per_chunk, [s[2] for s in spans]
    return merged


class AsrPipelineB200:
    """`pipeline("automatic-speech-recognition", model=..., chunk_length_s=15, batch_size=...)`-shaped callable for the
    scripts that reach the hot path through the HF ASR pipeline (run_speed_eval.py:53-59,76,
    run_short_form_eval.py:106-117,191):

        pipe = pipeline("automatic-speech-recognition", model=model, feature_extractor=fe, chunk_length_s=15)
        out = pipe(audio.copy(), generate_kwargs={"language": "ja", "task": "transcribe"})

    Input: a 1-D float waveform, a dict {"array"|"raw": waveform, "sampling_rate": sr} (datasets' audio column), or a
    list of those (-> list of results).  Output: {"text": str | None, "token_ids": [int]}; `text` needs a tokenizer
    (`tokenizer.decode(ids, skip_special_tokens=True)`), which cannot be fetched offline, so it is optional.
    Chunking, stride and the token-level LCS merge follow HF/pipelines/automatic_speech_recognition.py:61-84,428-444 and
    tokenization_whisper.py:1153-1270."""

    def __init__(self, model, feature_extractor=None, tokenizer=None, chunk_length_s: float = 0, stride_length_s=None,
                 batch_size: int = 1, device=None, torch_dtype=None, **unused):
        from .feature_extraction import WhisperFeatureExtractorB200
        self.model = model
        self.feature_extractor = feature_extractor or WhisperFeatureExtractorB200(
            feature_size=model.config.num_mel_bins, device=model.device)
        self.tokenizer = tokenizer
        self.chunk_length_s = chunk_length_s
        self.stride_length_s = stride_length_s
        self.batch_size = max(1, int(batch_size))

    def _one(self, item, generate_kwargs):
        sr = self.feature_extractor.sampling_rate
        if isinstance(item, dict):
            wav = item["array"] if "array" in item else item["raw"]
            in_sr = item.get("sampling_rate", sr)
            if in_sr != sr:
                raise ValueError(f"audio sampled at {in_sr} Hz but the feature extractor expects {sr} Hz "
                                 "(resampling needs torchaudio, which the HF pipeline also requires)")
        else:
            wav = item
        wav = np.asarray(wav, dtype=np.float32)
        if wav.ndim != 1:
            raise ValueError("We expect a single channel audio input for AutomaticSpeechRecognitionPipeline")
        gk = dict(generate_kwargs or {})
        lang, task = gk.pop("language", None), gk.pop("task", None)
        mnt = gk.pop("max_new_tokens", None)
        if gk:
            raise TypeError(f"unsupported generate_kwargs: {sorted(gk)}")
        chunk_s = self.chunk_length_s or self.feature_extractor.chunk_length
        if not self.chunk_length_s and wav.shape[0] > self.feature_extractor.n_samples:
            raise ValueError("audio longer than 30 s needs chunk_length_s (the sequential long-form path is "
                             "model.generate(..., return_timestamps=True))")
        ids = transcribe_longform(self.model, self.feature_extractor, wav, chunk_length_s=chunk_s,
                                  stride_length_s=self.stride_length_s, batch_size=self.batch_size, language=lang,
                                  task=task, max_new_tokens=mnt)
        text = self.tokenizer.decode(ids, skip_special_tokens=True) if self.tokenizer is not None else None
        return {"text": text, "token_ids": ids}

    def __call__(self, inputs, generate_kwargs=None, **kwargs):
        if kwargs.pop("return_timestamps", None):
            raise NotImplementedError("chunked pipeline with return_timestamps is outside the reference's path")
        if isinstance(inputs, (list, tuple)) and not (len(inputs) and np.isscalar(inputs[0])):
            return [self._one(x, generate_kwargs) for x in inputs]
        return self._one(inputs, generate_kwargs)


def pipeline(task: str = "automatic-speech-recognition", model=None, **kwargs) -> AsrPipelineB200:
    """Same call shape as `transformers.pipeline` for the one task the reference uses; `model` is a
    WhisperB200ForConditionalGeneration (or a checkpoint path, loaded with from_pretrained)."""
    if task != "automatic-speech-recognition":
        raise ValueError(f"only 'automatic-speech-recognition' is implemented, got {task!r}")
    if isinstance(model, str):
        from .modeling import WhisperB200ForConditionalGeneration
        import torch as _t
        model = WhisperB200ForConditionalGeneration.from_pretrained(
            model, torch_dtype=kwargs.get("torch_dtype") or _t.bfloat16, device=kwargs.get("device") or "cuda",
            max_batch=max(1, int(kwargs.get("batch_size", 1))), **(kwargs.pop("model_kwargs", None) or {}))
    kwargs.pop("model_kwargs", None)
    kwargs.pop("trust_remote_code", None)
    return AsrPipelineB200(model, **kwargs)
