"""Round-2 surface rows of SURVEY.md §8: teacher-forcing forward (f-4), generate(encoder_outputs=...) and batched
long-form pinned to HF goldens (f-3), the device-side chunker and the pipeline-shaped callable (f-2), the log-mel
producer loop (f-1), and the smaller pieces of the drop-in boundary (.pad, from_hf_model, the attention seam)."""
import numpy as np
import pytest
import torch

from _gpu_util import build_pair, state_dict_for
from _synth import KOTOBA, TEACHER, TINY, clips
from oracle.logmel_ref import logmel_batch_f64, logmel_f64

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def extra():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "tiny_extra.npz"))


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


# ---- f-4: teacher-forcing forward ------------------------------------------------------------------------------------
def test_teacher_forcing_logits_match_hf_golden(extra):
    model, ref = build_pair(TINY, torch.float32, max_batch=4)
    mel = torch.from_numpy(logmel_batch_f64(clips("UGS", 7), 128)).cuda()
    labels = torch.from_numpy(extra["tf_labels"])
    out = model(input_features=mel, labels=labels.cuda())
    assert out.logits.shape == (3, 24, TINY["vocab_size"]) and out.logits.dtype == torch.float32 and out.logits.is_cuda
    assert _rel(out.logits[:, :, ::53].cpu().numpy(), extra["tf_logits_sub"]) <= 1e-3
    assert abs(float(out.loss) - float(extra["tf_loss"])) <= 1e-3 * abs(float(extra["tf_loss"]))
    # encoder_outputs= + explicit decoder_input_ids (the shared-encoder branch of run_distillation.py:641-645)
    enc = model.get_encoder()(mel)
    out2 = model(encoder_outputs=enc, decoder_input_ids=labels.clamp(min=0).cuda())
    assert out2.loss is None
    assert _rel(out2.logits[:, :, ::53].cpu().numpy(), extra["tf_logits_ids_sub"]) <= 1e-3
    # against the oracle on the full vocabulary: teacher forcing == the cached decoder stepped position by position
    with torch.no_grad():
        enc_ref = ref.encode(mel.cpu())
        hid = ref.decode(labels.clamp(min=0), 0, [None] * TINY["decoder_layers"], ref.cross_kv(enc_ref))
        want = ref.logits(hid)
    assert _rel(out2.logits.cpu().numpy(), want.numpy()) <= 1e-3
    with pytest.raises(ValueError):
        model(input_features=mel)
    with pytest.raises(ValueError):
        model(input_features=mel, labels=torch.zeros((3, 449), dtype=torch.long))


def test_teacher_forcing_bf16_matches_cached_decode_steps():
    """bf16 tensor-core path at T = 128: the full-sequence forward agrees with the cached single-position decoder
    (itself pinned by the generate goldens) to bf16 rounding level at every position."""
    model, _ = build_pair(TINY, torch.bfloat16, max_batch=4)
    mel = torch.from_numpy(logmel_batch_f64(clips("UGS", 7), 128)).cuda()
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 50257, (3, 128), generator=g).cuda()
    out = model(input_features=mel, decoder_input_ids=ids).logits
    model.cross_kv(3)
    toks = ids.to(torch.int32)
    for pos in range(128):
        lg = model.step_logits(toks, pos)
        if pos in (0, 1, 31, 64, 127):
            assert ((out[:, pos] - lg).abs().max() / lg.abs().max()).item() <= 3e-2, pos


@pytest.mark.parametrize("arch,name,T", [(KOTOBA, "kotoba", 128), (TEACHER, "teacher", 128)])
def test_teacher_forcing_fullsize_fp32_matches_hf_golden(arch, name, T):
    """Full-size architectures at T = 128 (max_label_length of the reference's distillation configs) against logits
    HF computed on the CPU (tests/golden/make_golden.py teacher_tf)."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "teacher_tf.npz")
    if not os.path.exists(path):
        pytest.skip("teacher_tf.npz not generated")
    g = np.load(path)
    model, _ = build_pair(arch, torch.float32, max_batch=2)
    mel = torch.from_numpy(logmel_batch_f64(clips("GS", 3000), 128)).cuda()
    labels = torch.from_numpy(g["labels"])
    out = model(input_features=mel, labels=labels.cuda())
    assert _rel(out.logits[:, :, ::212].cpu().numpy(), g[f"{name}_logits_sub"]) <= 1e-3
    assert abs(float(out.loss) - float(g[f"{name}_loss"])) <= 1e-3 * abs(float(g[f"{name}_loss"]))


# ---- f-3: encoder_outputs and batched long-form pinned to HF ---------------------------------------------------------
def test_generate_from_encoder_outputs_matches_hf(extra):
    model, _ = build_pair(TINY, torch.float32, max_batch=4)
    mel = torch.from_numpy(logmel_batch_f64(clips("UGS", 7), 128)).cuda()
    enc = model.get_encoder()(mel)
    a = model.generate(encoder_outputs=enc, language="ja", task="transcribe", return_timestamps=False, max_length=40)
    assert np.array_equal(a.cpu().numpy(), extra["encout_ids_ts0"])
    from kotoba_whisper_b200.modeling import EncoderOutput
    for b in range(3):  # with timestamps HF keeps seeking over the same encoder output: several passes per row
        st = {}
        ids = model.generate(encoder_outputs=EncoderOutput(enc.last_hidden_state[b:b + 1]), language="ja",
                             task="transcribe", return_timestamps=True, max_length=40, stats=st)
        want = extra[f"encout_ids_ts1_row{b}"]
        assert ids.shape == want.shape and np.array_equal(ids.cpu().numpy(), want), (b, st)
    # a plain tensor is accepted too
    c = model.generate(encoder_outputs=enc.last_hidden_state, language="ja", task="transcribe",
                       return_timestamps=False, max_length=40)
    assert torch.equal(a, c)


def test_batched_longform_with_attention_mask_matches_hf(extra):
    """B = 3 recordings of 70 / 45 / 33 s in one batch: per-row max_frames from the attention mask, rows leaving the
    batch as they finish (HF _maybe_reduce_batch), segment slicing per row."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden import longform_batch
    model, _ = build_pair(TINY, torch.float32, max_batch=4)
    mel, mask = longform_batch()
    st = {}
    ids = model.generate(mel.cuda(), attention_mask=mask.cuda(), language="ja", task="transcribe",
                         return_timestamps=True, max_length=64, stats=st)
    want = extra["longform_b3_ids"]
    assert ids.shape == want.shape and np.array_equal(ids.cpu().numpy(), want), st
    with pytest.raises(ValueError):
        model.generate(mel.cuda(), language="ja", task="transcribe", return_timestamps=True, max_length=64)


def test_per_call_generation_config_must_match_baked_rules():
    from kotoba_whisper_b200 import WhisperB200GenerationConfig
    model, _ = build_pair(TINY, torch.float32, max_batch=2)
    mel = torch.from_numpy(logmel_batch_f64(clips("G", 3), 128)).cuda()
    ok = WhisperB200GenerationConfig()
    model.generate(mel, generation_config=ok, language="ja", task="transcribe", max_length=12)
    bad = WhisperB200GenerationConfig(suppress_tokens=(1, 2, 3))
    with pytest.raises(ValueError):
        model.generate(mel, generation_config=bad, language="ja", task="transcribe", max_length=12)
    with pytest.raises(NotImplementedError):
        model.generate(mel, prompt_ids=torch.tensor([1, 2, 3]), language="ja", task="transcribe", max_length=12)


# ---- boundary odds and ends --------------------------------------------------------------------------------------------
def test_from_hf_model_and_pad():
    from _hf import build_hf
    from kotoba_whisper_b200 import WhisperB200ForConditionalGeneration, WhisperFeatureExtractorB200
    hf = build_hf(TINY, seed=0)
    model = WhisperB200ForConditionalGeneration.from_hf_model(hf, max_batch=4, device="cuda:0")
    assert model.dtype == torch.float32 and model.config.d_model == TINY["d_model"]
    mel = torch.from_numpy(logmel_batch_f64(clips("UGS", 7), 128))
    with torch.no_grad():
        want = hf.generate(mel, language="ja", task="transcribe", return_timestamps=True, max_length=40, num_beams=1)
    got = model.generate(mel.cuda(), language="ja", task="transcribe", return_timestamps=True, max_length=40)
    assert torch.equal(got.cpu(), want)
    # from_pretrained with the reference's keywords (run_pseudo_labelling.py:224-232) on a locally saved checkpoint
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        hf.generation_config._from_model_config = False   # a real checkpoint ships its own generation_config.json
        hf.save_pretrained(d)
        m2 = WhisperB200ForConditionalGeneration.from_pretrained(d, torch_dtype=torch.float32,
                                                                 attn_implementation="sdpa", max_batch=4, device="cuda:0")
    got2 = m2.generate(mel.cuda(), language="ja", task="transcribe", return_timestamps=True, max_length=40)
    assert torch.equal(got2.cpu(), want)
    # .pad(): the collator's call (run_pseudo_labelling.py:154-158) — list of per-example features -> one batch
    fe = WhisperFeatureExtractorB200(feature_size=128, device="cuda:0")
    feats = fe(clips("GS", 11), sampling_rate=16000)["input_features"]
    batch = fe.pad([{"input_features": f} for f in feats], return_tensors="pt")
    assert batch["input_features"].shape == (2, 128, 3000) and batch.input_features.dtype == torch.float32
    assert np.array_equal(batch["input_features"].numpy(), np.stack(feats))
    from transformers import WhisperFeatureExtractor
    ref = WhisperFeatureExtractor(feature_size=128).pad([{"input_features": f} for f in feats], return_tensors="pt")
    assert torch.equal(ref["input_features"], batch["input_features"])
    # padding="longest" truncates at 30 s like HF
    long_clip = [np.zeros(16000 * 31, np.float32), np.zeros(16000 * 2, np.float32)]
    assert fe(long_clip, sampling_rate=16000, padding="longest")["input_features"].shape[-1] == 3000
    with pytest.raises(NotImplementedError):
        WhisperFeatureExtractorB200(feature_size=128, dither=0.1)


def test_attention_interface_seam_runs_inside_hf_encoder():
    """AttentionInterface.register("kwb200", fn): HF's own WhisperEncoder layer loop calling kw_attention
    (modeling_whisper.py:342-352) must reproduce its sdpa output."""
    from transformers import AttentionInterface
    from _hf import build_hf
    from kotoba_whisper_b200.attention_plugin import kwb200_attention_forward
    AttentionInterface.register("kwb200", kwb200_attention_forward)
    hf = build_hf(TINY, seed=0).cuda()
    mel = torch.from_numpy(logmel_batch_f64(clips("UG", 7), 128)).cuda()
    with torch.no_grad():
        want = hf.model.encoder(mel).last_hidden_state
        hf.config._attn_implementation = "kwb200"
        hf.model.encoder.config._attn_implementation = "kwb200"
        got = hf.model.encoder(mel).last_hidden_state
    assert ((got - want).abs().max() / want.abs().max()).item() <= 1e-3


# ---- f-2: device-side chunker + pipeline-shaped callable ---------------------------------------------------------------
def test_logmel_windows_equals_sliced_and_padded_clips():
    from kotoba_whisper_b200 import WhisperFeatureExtractorB200
    rng = np.random.default_rng(5)
    rec = (rng.standard_normal(16000 * 47 + 123) * 0.1).astype(np.float32)
    from kotoba_whisper_b200.pipeline import chunk_spans
    spans = list(chunk_spans(len(rec), 240000, 40000, 40000))
    assert len(spans) >= 4 and spans[-1][1] == len(rec)
    for nm in (80, 128):
        fe = WhisperFeatureExtractorB200(feature_size=nm, device="cuda:0")
        got = fe.logmel_windows(torch.from_numpy(rec).cuda(), [s for s, _, _, _ in spans],
                                [e - s for s, e, _, _ in spans]).cpu().numpy()
        want = np.stack([logmel_f64(rec[s:e], nm) for s, e, _, _ in spans])
        assert got.shape == want.shape == (len(spans), nm, 3000)
        assert np.abs(got - want).max() <= 1e-5
        host = fe([rec[s:e] for s, e, _, _ in spans], sampling_rate=16000)["input_features"]
        assert np.array_equal(got, host)  # same kernel, same arithmetic: bit-identical to the host-sliced path
    with pytest.raises(ValueError):
        fe.logmel_windows(torch.from_numpy(rec).cuda(), [len(rec) - 10], [240000])


def test_pipeline_callable_matches_host_chunked_path():
    from kotoba_whisper_b200 import WhisperFeatureExtractorB200, transcribe_longform
    from kotoba_whisper_b200.pipeline import pipeline
    model, _ = build_pair(TINY, torch.float32, max_batch=4)
    fe = WhisperFeatureExtractorB200(feature_size=128, device="cuda:0")
    rng = np.random.default_rng(17)
    audio = (rng.standard_normal(16000 * 50) * 0.1).astype(np.float32)
    pipe = pipeline("automatic-speech-recognition", model=model, feature_extractor=fe, chunk_length_s=15, batch_size=4)
    out = pipe(audio.copy(), generate_kwargs={"language": "ja", "task": "transcribe", "max_new_tokens": 20})
    st = {}
    want = transcribe_longform(model, fe, audio, chunk_length_s=15, batch_size=4, language="ja", task="transcribe",
                               max_new_tokens=20, device_chunker=False, stats=st)
    assert out["token_ids"] == want and out["text"] is None
    st2 = {}
    transcribe_longform(model, fe, audio, chunk_length_s=15, batch_size=4, language="ja", task="transcribe",
                        max_new_tokens=20, stats=st2)
    assert st2["h2d_bytes"] * 2 < st["h2d_bytes"]  # one upload of the recording vs every window padded to 30 s
    outs = pipe([{"array": audio[:16000 * 8], "sampling_rate": 16000}, {"raw": audio[:16000 * 20], "sampling_rate": 16000}],
                generate_kwargs={"language": "ja", "task": "transcribe", "max_new_tokens": 12})
    assert len(outs) == 2 and all(isinstance(o["token_ids"], list) for o in outs)
    with pytest.raises(ValueError):
        pipe({"array": audio, "sampling_rate": 8000})


# ---- f-1: producer loop ---------------------------------------------------------------------------------------------
def test_logmel_producer_streams_slabs_in_order():
    from kotoba_whisper_b200 import LogMelProducer, WhisperFeatureExtractorB200
    fe = WhisperFeatureExtractorB200(feature_size=80, device="cuda:0")
    all_clips = clips("GSUGSUGSUGSUG", 900) + [np.zeros(100, np.float32)]
    batches = [all_clips[i:i + 4] for i in range(0, len(all_clips), 4)]  # 4, 4, 4, 2
    prod = LogMelProducer(fe, slab_clips=4)
    seen = 0
    for first, feats in prod.produce(batches):
        assert first == seen
        want = logmel_batch_f64(all_clips[first:first + len(feats)], 80)
        assert feats.shape == want.shape and np.abs(feats - want).max() <= 1e-5
        seen += len(feats)
    assert seen == len(all_clips)
    assert prod.h2d_bytes == len(all_clips) * 480000 * 4 and prod.d2h_bytes == len(all_clips) * 80 * 3000 * 4


# ---- two batches in flight: GenerateStream == generate, batch by batch -------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("ts", [False, True])
def test_generate_stream_equals_generate_batch_by_batch(dtype, ts):
    from kotoba_whisper_b200 import WhisperFeatureExtractorB200
    model, _ = build_pair(TINY, dtype, max_batch=4)
    fe = WhisperFeatureExtractorB200(feature_size=128, device="cuda:0")
    batches = [fe(clips(fam, seed), sampling_rate=16000, return_tensors="pt", keep_on_device=True)["input_features"]
               for fam, seed in (("UGS", 11), ("GG", 12), ("SUGS", 13), ("U", 14))]   # ragged batch sizes 3, 2, 4, 1
    kw = dict(language="ja", task="transcribe", return_timestamps=ts, max_length=40)
    want = [model.generate(b, **kw).cpu() for b in batches]
    stats = {}
    stream = model.generate_stream(stats=stats, **kw)
    got = []
    for b in batches:
        r = stream.submit(b)
        if r is not None:
            got.append(r.cpu())
    while (r := stream.flush()) is not None:     # one batch per call, oldest first
        got.append(r.cpu())
    assert stream.flush() is None
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.shape == w.shape and torch.equal(g, w)
    # a second run through the same stream object (handles and captured graphs reused)
    assert stream.submit(batches[0]) is None
    again = [r.cpu() for r in (stream.submit(batches[1]), stream.flush(), stream.flush(), stream.flush()) if r is not None]
    assert len(again) == 2 and torch.equal(again[0], want[0]) and torch.equal(again[1], want[1])
    with pytest.raises(ValueError):
        stream.submit(torch.zeros((5, 128, 3000), device="cuda:0"))   # > max_batch
    with pytest.raises(ValueError):
        stream.submit(torch.zeros((1, 128, 6000), device="cuda:0"))   # long-form


# ---- coalesced decode batches: k submitted batches run as one device batch, ids unchanged ----------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("ts", [False, True])
def test_generate_stream_coalesce_equals_batch_by_batch(dtype, ts):
    from kotoba_whisper_b200 import WhisperFeatureExtractorB200
    model, _ = build_pair(TINY, dtype, max_batch=8)
    fe = WhisperFeatureExtractorB200(feature_size=128, device="cuda:0")
    batches = [fe(clips(fam, seed), sampling_rate=16000, return_tensors="pt", keep_on_device=True)["input_features"]
               for fam, seed in (("UGS", 21), ("GG", 22), ("SUGS", 23), ("U", 24), ("GU", 25))]   # ragged sizes, odd count
    kw = dict(language="ja", task="transcribe", return_timestamps=ts, max_length=40)
    want = [model.generate(b, **kw).cpu() for b in batches]
    stream = model.generate_stream(coalesce=2, **kw)
    got = []
    for b in batches:
        r = stream.submit(b)
        if r is not None:
            got.append(r.cpu())
    assert stream.device_batches == 2 and stream.buffered == 1
    while (r := stream.flush()) is not None:
        got.append(r.cpu())
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.shape == w.shape and torch.equal(g, w)
    with pytest.raises(ValueError):
        stream.submit(torch.zeros((9, 128, 3000), device="cuda:0"))   # > max_batch
    with pytest.raises(ValueError):
        model.generate_stream(coalesce=0, **kw)
    # a group that never fills (one batch, then flush) and `return_segments=True` (dict per submitted batch, as generate)
    stream = model.generate_stream(coalesce=3, return_segments=True, **kw)
    assert stream.submit(batches[2]) is None and stream.buffered == 1 and stream.device_batches == 0
    r = stream.flush()
    ref = model.generate(batches[2], return_segments=True, **kw)
    assert torch.equal(r["sequences"].cpu(), ref["sequences"].cpu()) and r["segments"] == ref["segments"]
    assert stream.flush() is None
    stream = model.generate_stream(coalesce=2, return_segments=True, **kw)
    got = [stream.submit(b) for b in batches[:4]]
    got = [r for r in got if r is not None]
    while (r := stream.flush()) is not None:
        got.append(r)
    assert len(got) == 4
    for r, b in zip(got, batches[:4]):
        ref = model.generate(b, return_segments=True, **kw)
        assert torch.equal(r["sequences"].cpu(), ref["sequences"].cpu()) and r["segments"] == ref["segments"]


def test_kotoba_bf16_coalesced_128_row_decode_is_row_identical():
    """The benchmarked schedule: two 64-utterance batches decoded as one 128-row device batch (decode-time GEMMs with the
    batch as a 128-wide MMA N, fused vocabulary epilogue over 128 rows).  Every utterance's ids must be the bits a
    64-row `generate` gives: rows are independent in every kernel and the k order of a row's dot products is the same."""
    from kotoba_whisper_b200 import WhisperB200ForConditionalGeneration, WhisperFeatureExtractorB200
    from _gpu_util import state_dict_for
    from _synth import KOTOBA, clip
    sd, cfg = state_dict_for(tuple(sorted(KOTOBA.items())))
    model = WhisperB200ForConditionalGeneration.from_state_dict(sd, cfg, dtype=torch.bfloat16, max_batch=128, device="cuda:0")
    fe = WhisperFeatureExtractorB200(feature_size=128, device="cuda:0")
    audio = [clip("UGSG"[i % 4], 9000 + i) for i in range(128)]
    halves = [fe(audio[i:i + 64], sampling_rate=16000, return_tensors="pt", keep_on_device=True)["input_features"]
              for i in (0, 64)]
    for ts in (False, True):
        kw = dict(language="ja", task="transcribe", return_timestamps=ts, max_length=64)
        want = [model.generate(h, **kw).cpu() for h in halves]
        both = model.generate(torch.cat(halves), **kw).cpu()          # plain 128-row generate
        pad = model.generation_config.pad_token_id
        for i, w in enumerate(want):
            g = both[64 * i: 64 * (i + 1)]
            L = max(w.shape[1], 1)
            assert torch.equal(g[:, :w.shape[1]], w) and bool((g[:, w.shape[1]:] == pad).all()), (ts, i, L)
        stream = model.generate_stream(coalesce=2, **kw)
        assert stream.submit(halves[0]) is None and stream.submit(halves[1]) is None
        got = []
        while (r := stream.flush()) is not None:
            got.append(r.cpu())
        assert len(got) == 2 and all(torch.equal(g, w) for g, w in zip(got, want)), ts
