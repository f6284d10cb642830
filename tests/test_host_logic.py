"""Host-side logic of the product (no GPU): segment slicing, prompt construction, chunking and the token merge,
checked against the oracle and against transformers' own helpers."""
import numpy as np
import pytest
import torch

from kotoba_whisper_b200 import modeling as km
from kotoba_whisper_b200.pipeline import chunk_spans, merge_chunk_tokens
from oracle.whisper_ref import ArchConfig, GenConfig, WhisperRef


class _Host(km.WhisperB200ForConditionalGeneration):
    def __init__(self):  # host logic only: no device, no library
        self.generation_config = km.WhisperB200GenerationConfig()
        self.config = km.WhisperB200Config()


def _ref():
    r = WhisperRef.__new__(WhisperRef)
    r.gen, r.arch = GenConfig(), ArchConfig()
    return r


def test_split_segments_matches_oracle_and_hf():
    from transformers.models.whisper.generation_whisper import WhisperGenerationMixin
    host, ref = _Host(), _ref()
    tb = 50365
    rng = np.random.default_rng(0)
    for trial in range(300):
        n = int(rng.integers(1, 40))
        seq = []
        for _ in range(n):
            seq.append(int(tb + rng.integers(0, 1500)) if rng.random() < 0.4 else int(rng.integers(0, 50000)))
        nf = int(rng.integers(20, 3001))
        a = host._split_segments(list(seq), nf)
        b = ref.retrieve_segment(list(seq), nf)
        assert a == b
        segs, off = WhisperGenerationMixin._retrieve_segment(
            seek_sequence=torch.tensor(seq), seek_outputs=[None], time_offset=torch.zeros(1, dtype=torch.float64),
            timestamp_begin=tb, seek_num_frames=torch.tensor([nf]), time_precision=0.02, time_precision_features=0.01,
            input_stride=2, prev_idx=0, idx=0, return_token_timestamps=False, decoder_input_ids=torch.zeros(1, 3))
        assert [s["tokens"].tolist() for s in segs] == a[0]
        assert int(off) == a[1]


def test_init_tokens():
    host, ref = _Host(), _ref()
    assert host._init_tokens("ja", "transcribe", True) == [50258, 50266, 50360]
    assert host._init_tokens("ja", "transcribe", False) == [50258, 50266, 50360, 50364]
    assert host._init_tokens("en", "translate", True) == [50258, 50259, 50359]
    assert host._init_tokens("<|ja|>", None, False) == [50258, 50266, 50360, 50364]
    assert host._init_tokens("japanese", None, True) == [50258, 50266, 50360]
    for lang, task, ts in (("ja", "transcribe", True), ("en", "translate", False), ("ja", None, False)):
        assert host._init_tokens(lang, task, ts) == ref.init_tokens(lang, task, ts)
    with pytest.raises(ValueError):
        host._init_tokens("xx", "transcribe", True)
    with pytest.raises(ValueError):
        host._init_tokens("ja", "summarise", True)


def test_merge_matches_hf_lcs():
    from transformers.models.whisper.tokenization_whisper import _find_longest_common_sequence
    rng = np.random.default_rng(1)
    for trial in range(200):
        base = rng.integers(0, 50, size=int(rng.integers(30, 120))).tolist()
        seqs, pos = [], 0
        while pos < len(base):
            n = int(rng.integers(8, 25))
            chunk = base[max(0, pos - int(rng.integers(0, 6))): pos + n]
            if rng.random() < 0.3 and len(chunk) > 3:  # recognition noise inside the overlap
                chunk[int(rng.integers(0, len(chunk)))] = int(rng.integers(50, 60))
            seqs.append(chunk)
            pos += n
        assert merge_chunk_tokens(seqs) == _find_longest_common_sequence(seqs)
    assert merge_chunk_tokens([]) == []
    assert merge_chunk_tokens([[1, 2, 3]]) == [1, 2, 3]


def test_chunk_spans_match_hf_chunk_iter():
    from transformers.pipelines.automatic_speech_recognition import chunk_iter as hf_chunk_iter

    class FakeFE:
        sampling_rate = 16000

        def __call__(self, chunk, **kw):
            return {"n": len(chunk)}

    for n in (1000, 240000, 240001, 400000, 16000 * 300, 16000 * 3600 + 123, 160000, 200000):
        x = np.zeros(n, np.float32)
        want = [(d["stride"], d["is_last"], d["n"]) for d in hf_chunk_iter(x, FakeFE(), 240000, 40000, 40000)]
        got = [(st, last, e - s) for s, e, st, last in chunk_spans(n, 240000, 40000, 40000)]
        assert got == want
    assert len(list(chunk_spans(16000 * 3600, 240000, 40000, 40000))) == 360  # cfg4: 1 h -> 360 windows


def test_shard_range_is_a_partition():
    from kotoba_whisper_b200.distributed import shard_range
    for n in (0, 1, 7, 64, 257):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_stripped_lengths_match_per_row_rule():
    """Vectorised pad / eos stripping == the per-row rule of generation_whisper.py:1063-1086."""
    from kotoba_whisper_b200.modeling import WhisperB200ForConditionalGeneration as M
    rng = np.random.default_rng(5)
    for pad, eos in ((50257, 50257), (50256, 50257)):
        rows = []
        for _ in range(400):
            T = 12
            n = int(rng.integers(0, T + 1))
            seq = rng.integers(0, 300, size=T)
            if rng.random() < 0.7 and n < T:
                seq[n] = eos
                seq[n + 1:] = pad
            if rng.random() < 0.2:
                seq[int(rng.integers(0, T))] = pad   # a stray pad id in the middle
            rows.append(seq)
        arr = np.stack(rows)
        got = M._stripped_lengths(arr, pad, eos)
        for r, g in zip(rows, got):
            seq = r.tolist()
            if seq[-1] == pad:
                n_pad = sum(1 for t in seq if t == pad) - (1 if pad == eos else 0)
                if n_pad > 0:
                    seq = seq[:-n_pad]
            if seq and seq[-1] == eos:
                seq = seq[:-1]
            assert len(seq) == int(g)
