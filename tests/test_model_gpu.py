"""End-to-end parity of the CUDA path with the oracle and the HF goldens.

Bars (BASELINE.json north_star): fp32 — encoder hidden states and decoder logits within 1e-3 relative, greedy token ids
bit-identical; bf16 — encoder cosine >= 0.999 and token sequences compared against the fp32 oracle on bf16-rounded
weights (near-tie flips triaged by the oracle's top-2 margin)."""
import numpy as np
import pytest
import torch

from _gpu_util import build_pair
from _synth import KOTOBA, TEACHER, TINY, TINY80, clips
from oracle.logmel_ref import logmel_batch_f64

pytestmark = pytest.mark.gpu
CASES = [(True, 40, "ja", "transcribe"), (True, 128, "ja", "transcribe"), (False, 40, "ja", "transcribe"),
         (False, 128, "ja", "transcribe"), (True, 64, "en", "translate")]


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("name,arch", [("tiny", TINY), ("tiny80", TINY80)])
def test_tiny_fp32_matches_oracle_and_goldens(golden, name, arch):
    model, ref = build_pair(arch, torch.float32, max_batch=4)
    mel = torch.from_numpy(logmel_batch_f64(clips("UGS", 7), arch["num_mel_bins"]))
    enc = model.encode(mel.cuda()).cpu()
    with torch.no_grad():
        enc_ref = ref.encode(mel)
    assert _rel(enc, enc_ref) <= 1e-3
    assert np.abs(enc[:, ::97, ::5].numpy() - golden["tiny"][f"{name}_enc_sub"]).max() <= 1e-3 * enc_ref.abs().max()
    # raw logits of the prompt positions and one generated position
    model.cross_kv(3)
    toks = torch.tensor([[50258, 50266, 50360, 50365, 11]] * 3, dtype=torch.int32, device="cuda")
    with torch.no_grad():
        cross = ref.cross_kv(enc_ref)
        cache = [None] * arch["decoder_layers"]
        hid = ref.decode(toks.long().cpu(), 0, cache, cross)
        lg_ref = ref.logits(hid)
    for pos in range(5):
        lg = model.step_logits(toks, pos).cpu()
        assert _rel(lg, lg_ref[:, pos]) <= 1e-3, pos
    for ts, ml, lang, task in CASES:
        ids = model.generate(mel.cuda(), language=lang, task=task, return_timestamps=ts, max_length=ml)
        want = golden["tiny"][f"{name}_ids_ts{int(ts)}_ml{ml}_{lang}_{task}"]
        assert ids.is_cuda and ids.dtype == torch.long
        assert ids.shape == want.shape and np.array_equal(ids.cpu().numpy(), want), (name, ts, ml, lang)
    # max_new_tokens, batch > max_batch chunking, host input
    big = torch.from_numpy(logmel_batch_f64(clips("GSUGSUG", 50), arch["num_mel_bins"]))
    a = model.generate(big, language="ja", task="transcribe", return_timestamps=True, max_new_tokens=20)
    with torch.no_grad():
        b = ref.generate(big, language="ja", task="transcribe", return_timestamps=True, max_length=23)
    assert not a.is_cuda and torch.equal(a, b)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_decoder_self_attention_over_long_histories(dtype):
    """Raw logits at positions 0..299 (teacher-forced random history) against the oracle: the decode-time self-attention
    walks its cache in 128-row chunks, so positions past 128 and 256 exercise the multi-chunk path."""
    model, ref = build_pair(TINY, dtype, max_batch=4, oracle_weights="rounded" if dtype == torch.bfloat16 else "same")
    mel = torch.from_numpy(logmel_batch_f64(clips("UG", 31), 128))
    if dtype == torch.bfloat16:
        mel = mel.to(torch.bfloat16).to(torch.float32)
    model.encode(mel.cuda())
    model.cross_kv(2)
    g = torch.Generator().manual_seed(3)
    toks = torch.randint(0, 50257, (2, 300), generator=g, dtype=torch.int32)
    toks[:, 0] = 50258
    with torch.no_grad():
        enc_ref = ref.encode(mel)
        cross = ref.cross_kv(enc_ref)
        lg_ref = ref.logits(ref.decode(toks.long(), 0, [None] * TINY["decoder_layers"], cross))
    tol = 1e-3 if dtype == torch.float32 else 5e-2   # bf16: rounding noise; an indexing bug is an O(1) error
    td = toks.cuda()
    for pos in range(300):
        lg = model.step_logits(td, pos)
        if pos in (0, 1, 31, 127, 128, 129, 200, 255, 256, 299):
            assert _rel(lg.cpu(), lg_ref[:, pos]) <= tol, (pos, _rel(lg.cpu(), lg_ref[:, pos]))


def test_tiny_errors_and_encoder_outputs():
    model, ref = build_pair(TINY, torch.float32, max_batch=4)
    mel = torch.from_numpy(logmel_batch_f64(clips("GS", 3), 128)).cuda()
    with pytest.raises(ValueError):
        model.encode(mel[:, :, :2000])
    with pytest.raises(ValueError):
        model.generate(mel, language="xx", task="transcribe")
    with pytest.raises(ValueError):
        model.generate(mel, language="ja", task="transcribe", max_new_tokens=446)
    with pytest.raises(NotImplementedError):
        model.generate(mel, language="ja", task="transcribe", num_beams=5)
    enc = model.get_encoder()(mel)
    a = model.generate(encoder_outputs=enc, language="ja", task="transcribe", return_timestamps=False, max_length=40)
    b = model.generate(mel, language="ja", task="transcribe", return_timestamps=False, max_length=40)
    # encoder_outputs -> single pass; compare the first pass' tokens
    n = min(a.shape[1], b.shape[1])
    assert torch.equal(a[:, : min(n, 20)], b[:, : min(n, 20)])


def test_tiny_longform_seek_loop():
    """> 30 s input with timestamps: the seek loop, ragged `attention_mask` lengths and batch shrinking."""
    model, ref = build_pair(TINY, torch.float32, max_batch=4)
    from oracle.logmel_ref import logmel_f64
    rng = np.random.default_rng(9)
    audio = [(rng.standard_normal(16000 * 70) * 0.1).astype(np.float32)]
    mel = torch.from_numpy(np.stack([logmel_f64(a, 128, n_samples=16000 * 70) for a in audio]))
    a = model.generate(mel.cuda(), language="ja", task="transcribe", return_timestamps=True, max_length=64)
    with torch.no_grad():
        b = ref.generate(mel, language="ja", task="transcribe", return_timestamps=True, max_length=64)
    assert torch.equal(a.cpu(), b)
    with pytest.raises(ValueError):
        model.generate(mel.cuda(), language="ja", task="transcribe", return_timestamps=False)


def _first_divergence(a, b):
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return i
    return n if len(a) != len(b) else -1


def test_tiny_bf16_vs_rounded_weight_oracle():
    """bf16 bar on the tiny architecture: encoder cosine >= 0.999 vs the fp32 oracle on bf16-rounded weights; the
    product's fp32 path on those weights is token-identical to the oracle; the bf16 path agrees with it on every
    utterance up to near-ties (divergences whose fp32 logit gap is below 4 sigma of the bf16 logit noise)."""
    from kotoba_whisper_b200 import WhisperB200ForConditionalGeneration
    from kotoba_whisper_b200.parity import bf16_token_parity, rounded_state_dict
    from _gpu_util import state_dict_for
    model, ref = build_pair(TINY, torch.bfloat16, max_batch=12, oracle_weights="rounded")
    sd, cfg = state_dict_for(tuple(sorted(TINY.items())))
    m32 = WhisperB200ForConditionalGeneration.from_state_dict(rounded_state_dict(sd), cfg, dtype=torch.float32,
                                                              max_batch=12, device="cuda:0")
    mel = torch.from_numpy(logmel_batch_f64(clips("UGSUGSUGSUGS", 7), 128))
    mel_r = mel.to(torch.bfloat16).to(torch.float32)
    enc = model.encode(mel.cuda()).cpu()
    with torch.no_grad():
        enc_ref = ref.encode(mel_r)
    cos = torch.nn.functional.cosine_similarity(enc.flatten(1), enc_ref.flatten(1), dim=1)
    assert cos.min() >= 0.999, cos
    ids32 = m32.generate(mel_r.cuda(), language="ja", task="transcribe", return_timestamps=False, max_length=64).cpu()
    with torch.no_grad():
        want = ref.generate(mel_r, language="ja", task="transcribe", return_timestamps=False, max_length=64)
    assert torch.equal(ids32, want), "fp32 path on rounded weights must be token-identical to the oracle"
    res = bf16_token_parity(model, m32, mel, max_length=64)
    assert res["adjusted_identical"] == res["utterances"], res
    assert res["raw_identical"] >= 1, res


def _check_fullsize(golden, name, arch, spec, seed, cases, max_batch):
    model, _ = build_pair(arch, torch.float32, max_batch=max_batch)
    g = golden[name]
    mel = torch.from_numpy(logmel_batch_f64(clips(spec, seed), arch["num_mel_bins"])).cuda()
    enc = model.encode(mel).cpu()
    sub = enc[:, ::97, ::5].numpy()
    scale = np.abs(g[f"{name}_enc_sub"]).max()
    assert np.abs(sub - g[f"{name}_enc_sub"]).max() <= 1e-3 * scale
    assert np.abs(enc.abs().mean(dim=(1, 2)).numpy() - g[f"{name}_enc_absmean"]).max() <= 1e-4
    model.cross_kv(mel.shape[0])
    toks = torch.tensor([[50258, 50266, 50360, 50364]] * mel.shape[0], dtype=torch.int32, device="cuda")
    for pos in range(3):
        model.step_logits(toks, pos)
    lg0 = model.step_logits(toks, 3).cpu().numpy()[:, ::53]
    want0 = g[f"{name}_logits0_sub"]
    assert np.abs(lg0 - want0).max() <= 1e-3 * np.abs(want0).max()
    for ts, ml in cases:
        st = {}
        ids = model.generate(mel, language="ja", task="transcribe", return_timestamps=ts, max_length=ml, stats=st).cpu().numpy()
        want = g[f"{name}_ids_ts{int(ts)}_ml{ml}_ja_transcribe"]
        if ids.shape != want.shape or not np.array_equal(ids, want):
            rows = [(b, _first_divergence(ids[b].tolist(), want[b].tolist())) for b in range(min(len(ids), len(want)))]
            pytest.fail(f"{name} ts={ts}: token mismatch, first divergences {rows}, shapes {ids.shape} vs {want.shape}, "
                        f"passes {st}")


def test_kotoba_fp32_tokens_bit_identical(golden):
    """BASELINE configs[0]: kotoba-whisper-v2.0 architecture, fp32, batch 4 x 30 s, ja/transcribe."""
    _check_fullsize(golden, "kotoba", KOTOBA, "UGSG", 1000, [(True, 128), (False, 128)], max_batch=4)


def test_teacher_fp32_tokens_bit_identical(golden):
    """configs[2] architecture (32 decoder layers) at the CPU-runnable batch the golden was taken on."""
    _check_fullsize(golden, "teacher", TEACHER, "GS", 3000, [(True, 128)], max_batch=2)


@pytest.mark.parametrize("arch", [TINY, TINY80])
def test_bf16_tensor_path_matches_simt_path(arch):
    """Same bf16 model through the tcgen05 kernels (GEMM, skinny GEMM with PDL, attention) and through the SIMT kernels:
    encoder output, per-position logits and greedy tokens must agree to bf16-rounding level."""
    from kotoba_whisper_b200 import _lib
    lib = _lib.load()
    model, _ = build_pair(arch, torch.bfloat16, max_batch=4)
    mel = torch.from_numpy(logmel_batch_f64(clips("UGS", 7), arch["num_mel_bins"])).cuda()
    toks = torch.tensor([[50258, 50266, 50360, 50365, 11, 12]] * 3, dtype=torch.int32, device="cuda")
    res = {}
    for impl in (0, 1):
        lib.kw_set_gemm_impl(impl)
        enc = model.encode(mel)
        model.cross_kv(3)
        lg = torch.stack([model.step_logits(toks, pos) for pos in range(6)])
        ids = model.generate(mel, language="ja", task="transcribe", return_timestamps=True, max_length=48)
        res[impl] = (enc.clone(), lg.clone(), ids.clone())
    lib.kw_set_gemm_impl(0)
    enc_t, lg_t, ids_t = res[0]
    enc_s, lg_s, ids_s = res[1]
    assert _rel(enc_t, enc_s) <= 3e-2
    cos = torch.nn.functional.cosine_similarity(enc_t.flatten(1), enc_s.flatten(1), dim=1)
    assert cos.min() >= 0.9995
    assert _rel(lg_t, lg_s) <= 3e-2
    # one raw greedy pass per path: identical, or first divergence at a near-tie of the SIMT path's own logits (gap
    # below 4 sigma of the tensor-vs-SIMT logit difference)
    sigma = (lg_t - lg_s).double().pow(2).mean().sqrt().item()
    prompt = [50258, 50266, 50360, 50364]
    raw = {}
    for impl in (0, 1):
        lib.kw_set_gemm_impl(impl)
        model.encode(mel, return_hidden=False)
        raw[impl] = model._greedy_pass(3, prompt, 48, False)[:, 4:]
    try:  # impl 1 (SIMT) is still selected and holds the encoder state
        model.cross_kv(3)
        for b in range(3):
            a, c = raw[0][b].tolist(), raw[1][b].tolist()
            j = next((i for i in range(len(a)) if a[i] != c[i]), -1)
            if j < 0:
                continue
            hist = torch.tensor([prompt + c[:j]] * 3, dtype=torch.int32, device="cuda")
            for pos in range(hist.shape[1]):
                lg = model.step_logits(hist, pos)
            gap = float(lg[b, c[j]] - lg[b, a[j]])
            assert abs(gap) < 4 * sigma, (b, j, gap, sigma)
    finally:
        lib.kw_set_gemm_impl(0)


def test_kotoba_bf16_token_parity_bar():
    """BASELINE north_star bf16 bar on the benchmarked configuration (kotoba-v2.0 architecture, greedy short-form,
    max_length 128), 128 utterances: encoder cosine >= 0.999 and token sequences identical on >= 99 % of utterances
    against the exact-fp32 CUDA path (itself bit-identical to HF, test_kotoba_fp32_tokens_bit_identical) on the SAME
    bf16-rounded weights and features.  With random-init weights the top-2 logit gap is of the order of the bf16 logit
    noise, so identity is counted up to near-ties (kotoba_whisper_b200/parity.py: a first divergence whose fp32 logit gap
    is below tau = 4 sigma of the measured bf16 logit noise is a tie); the raw figure must not be worse than what HF's
    own bf16 (sdpa, same GPU, same weights) achieves against the same fp32 tokens, within 10 points of sampling noise.
    tools/bf16_parity.py writes the same measurement to profiles/r2_bf16_parity.json."""
    from kotoba_whisper_b200 import WhisperB200ForConditionalGeneration, WhisperFeatureExtractorB200
    from kotoba_whisper_b200.parity import bf16_token_parity, rounded_state_dict, first_divergence, _trim
    from _gpu_util import state_dict_for
    from _hf import build_hf
    from _synth import clip
    N = 128
    sd, cfg = state_dict_for(tuple(sorted(KOTOBA.items())))
    rounded = rounded_state_dict(sd)
    m16 = WhisperB200ForConditionalGeneration.from_state_dict(rounded, cfg, dtype=torch.bfloat16, max_batch=32, device="cuda:0")
    m32 = WhisperB200ForConditionalGeneration.from_state_dict(rounded, cfg, dtype=torch.float32, max_batch=32, device="cuda:0")
    fe = WhisperFeatureExtractorB200(feature_size=128, device="cuda:0")
    audio = [clip("UGSG"[i % 4], 7000 + i) for i in range(N)]
    mel = torch.cat([fe(audio[i:i + 32], sampling_rate=16000, return_tensors="pt", keep_on_device=True)["input_features"]
                     for i in range(0, N, 32)])
    mel_r = mel.to(torch.bfloat16).to(torch.float32)
    e16, e32 = m16.encode(mel_r[:8]), m32.encode(mel_r[:8])
    cos = torch.nn.functional.cosine_similarity(e16.flatten(1), e32.flatten(1), dim=1)
    assert cos.min().item() >= 0.999, cos
    res = bf16_token_parity(m16, m32, mel, max_length=128)
    assert res["adjusted_pct"] >= 99.0, res
    # HF's own bf16 against the same fp32 tokens
    hf = build_hf(KOTOBA, seed=0)
    hf.load_state_dict(rounded)
    hf = hf.to(torch.bfloat16).cuda()
    pad = m32.generation_config.pad_token_id
    same_hf = 0
    with torch.no_grad():
        for i in range(0, N, 32):
            x = mel_r[i:i + 32]
            want = [_trim(r, pad) for r in m32.generate(x, language="ja", task="transcribe", return_timestamps=False,
                                                        max_length=128).cpu().tolist()]
            got = [_trim(r, pad) for r in hf.generate(x.to(torch.bfloat16), language="ja", task="transcribe",
                                                      return_timestamps=False, max_length=128, num_beams=1).cpu().tolist()]
            same_hf += sum(int(first_divergence(p, q) < 0) for p, q in zip(got, want))
    assert res["raw_identical"] >= same_hf - int(0.10 * N), (res["raw_identical"], same_hf)
