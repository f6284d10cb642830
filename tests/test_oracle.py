"""The oracle (oracle/) against the golden fixtures produced by the reference implementation (transformers' Whisper,
tests/golden/make_golden.py) and, when importable, against transformers live.  CPU only."""
import numpy as np
import pytest
import torch

from _synth import KOTOBA, TEACHER, TINY, TINY80, clips
from oracle.logmel_ref import frame_attention_mask, logmel_batch_f64, logmel_f64, mel_filter_bank
from oracle.whisper_ref import ArchConfig, GenConfig, WhisperRef


def _logmel_inputs(nm):
    cl = clips("UGSS", 500 + nm)
    cl.append(np.zeros(1000, np.float32))
    return cl


@pytest.mark.parametrize("nm", [80, 128])
def test_logmel_oracle_matches_hf_golden(golden, nm):
    g = golden["logmel"]
    cl = _logmel_inputs(nm)
    ref = logmel_batch_f64(cl, nm)
    assert ref.shape == (5, nm, 3000) and ref.dtype == np.float32
    # float64 truth vs HF's float64 numpy path: cast-level agreement; vs HF's fp32 torch path: HF's own 1e-5-class gap
    assert np.abs(ref[:, :, ::37] - g[f"hf_numpy_{nm}"]).max() <= 5e-7
    assert np.abs(ref[:, :, :8] - g[f"hf_numpy_head_{nm}"]).max() <= 5e-7
    assert np.abs(ref[:, :, -8:] - g[f"hf_numpy_tail_{nm}"]).max() <= 5e-7
    assert np.abs(ref[:, :, ::37] - g[f"hf_torch_{nm}"]).max() <= 5e-5
    assert np.abs(ref.astype(np.float64).sum(-1) - g[f"hf_numpy_rowsum_{nm}"]).max() <= 2e-3
    lens = g[f"clip_len_{nm}"]
    assert (frame_attention_mask(lens).sum(-1) == g[f"mask_sum_{nm}"]).all()


def test_logmel_known_answers():
    # SURVEY.md §8c known-answer facts
    for nm, nnz in ((80, 391), (128, 394)):
        fb = mel_filter_bank(nm)
        assert fb.shape == (201, nm) and fb.dtype == np.float64
        assert int((fb != 0).sum()) == nnz
        assert not fb[0].any() and not fb[200].any()
    five = np.random.default_rng(0).standard_normal(5 * 16000).astype(np.float32) * 0.1
    out = logmel_f64(five, 128)
    assert out.shape == (128, 3000)
    assert np.all(out[:, 600:] == out[0, -1])  # zero-padded region is one constant: (max(-10, clipmax-8)+4)/4
    assert frame_attention_mask([5 * 16000]).sum() == 500
    silent = logmel_f64(np.zeros(16000, np.float32), 80)
    assert np.all(silent == -1.5)
    # batched == single clip (per-clip max)
    a = logmel_batch_f64([five, five * 3], 80)
    assert np.array_equal(a[0], logmel_f64(five, 80))


CASES = [(True, 40, "ja", "transcribe"), (True, 128, "ja", "transcribe"), (False, 40, "ja", "transcribe"),
         (False, 128, "ja", "transcribe"), (True, 64, "en", "translate")]


def _hf_state(arch):
    from _hf import build_hf
    return build_hf(arch)


@pytest.mark.parametrize("name,arch", [("tiny", TINY), ("tiny80", TINY80)])
def test_whisper_oracle_matches_hf_golden_tiny(golden, name, arch):
    g = golden["tiny"]
    hf = _hf_state(arch)  # only for the seeded weights; arithmetic below is the oracle's
    ref = WhisperRef(hf.state_dict(), ArchConfig(**arch))
    mel = torch.from_numpy(logmel_batch_f64(clips("UGS", 7), arch["num_mel_bins"]))
    with torch.no_grad():
        enc = ref.encode(mel)
        sub = enc[:, ::97, ::5].numpy()
        assert np.abs(sub - g[f"{name}_enc_sub"]).max() <= 1e-4
        for ts, ml, lang, task in CASES:
            ids = ref.generate(mel, language=lang, task=task, return_timestamps=ts, max_length=ml)
            want = g[f"{name}_ids_ts{int(ts)}_ml{ml}_{lang}_{task}"]
            assert ids.shape == want.shape and np.array_equal(ids.numpy(), want), (name, ts, ml)


def test_whisper_oracle_matches_hf_live():
    hf = _hf_state(TINY)
    ref = WhisperRef(hf.state_dict(), ArchConfig(**TINY))
    mel = torch.from_numpy(logmel_batch_f64(clips("SG", 91), 128))
    with torch.no_grad():
        assert (ref.encode(mel) - hf.model.encoder(mel).last_hidden_state).abs().max() < 1e-5
        a = hf.generate(mel, language="ja", task="transcribe", return_timestamps=True, max_length=96, num_beams=1)
        b = ref.generate(mel, language="ja", task="transcribe", return_timestamps=True, max_length=96)
        assert torch.equal(a, b)
        # one raw decoder step
        enc = ref.encode(mel)
        cross = ref.cross_kv(enc)
        cache = [None] * TINY["decoder_layers"]
        ids = torch.tensor([[50258, 50266, 50360]] * 2)
        lg = ref.logits(ref.decode(ids, 0, cache, cross)[:, -1])
        lg_hf = hf(input_features=mel, decoder_input_ids=ids).logits[:, -1]
        assert (lg - lg_hf).abs().max() < 1e-4


def test_generation_config_constants():
    from transformers.models.whisper.configuration_whisper import NON_SPEECH_TOKENS_MULTI
    from transformers.models.whisper.tokenization_whisper import LANGUAGES
    import oracle.whisper_ref as wr
    import kotoba_whisper_b200.modeling as km
    assert wr._NON_SPEECH_82 == NON_SPEECH_TOKENS_MULTI[:82]
    assert list(LANGUAGES) == wr._LANG_CODES == km.LANGUAGE_CODES
    assert list(GenConfig().suppress_tokens) == km._V3_SUPPRESS
    assert GenConfig().lang_to_id["<|ja|>"] == 50266 and GenConfig().timestamp_begin == 50365


@pytest.mark.slow
@pytest.mark.parametrize("name,arch,spec,seed", [("kotoba", KOTOBA, "UGSG", 1000), ("teacher", TEACHER, "GS", 3000)])
def test_whisper_oracle_matches_hf_golden_fullsize(golden, name, arch, spec, seed):
    g = golden[name]
    hf = _hf_state(arch)
    ref = WhisperRef(hf.state_dict(), ArchConfig(**arch))
    del hf
    mel = torch.from_numpy(logmel_batch_f64(clips(spec, seed), arch["num_mel_bins"]))
    with torch.no_grad():
        ids = ref.generate(mel, language="ja", task="transcribe", return_timestamps=True, max_length=128)
    assert np.array_equal(ids.numpy(), g[f"{name}_ids_ts1_ml128_ja_transcribe"])


def test_chunking_oracle_matches_hf():
    from transformers.models.whisper.tokenization_whisper import _find_longest_common_sequence
    from transformers.pipelines.automatic_speech_recognition import chunk_iter
    from oracle.chunking_ref import chunk_bounds, longest_common_sequence_merge

    class FakeFE:
        sampling_rate = 16000

        def __call__(self, chunk, **kw):
            return {"n": len(chunk)}

    for n in (1000, 240000, 400001, 16000 * 100):
        want = [d["stride"] for d in chunk_iter(np.zeros(n, np.float32), FakeFE(), 240000, 40000, 40000)]
        assert [s for _, _, s in chunk_bounds(n, 240000, 40000, 40000)] == want
    rng = np.random.default_rng(4)
    for _ in range(50):
        seqs = [rng.integers(0, 12, size=int(rng.integers(3, 30))).tolist() for _ in range(int(rng.integers(1, 5)))]
        assert longest_common_sequence_merge(seqs) == _find_longest_common_sequence(seqs)
