"""Throughput of the other BASELINE.json configs (parity-test cases, not bench lines) on one B200; prints ONE JSON object
(kept as profiles/r2_configs.json):
   cfg1 fp32 kotoba B=4 timestamps; cfg3 teacher bf16 B=32 timestamps; cfg4 long-form 1 h in 15 s chunks (device-side
   chunker vs host path); cfg5 log-mel sweep 80/128 mels x 1k..64k clips (resident slabs and the producer loop with
   H2D + D2H inside the timing); lat = batch-1 pipeline latency for 10/30/60/300 s clips (the reference's own published
   measurement, eval_pipeline/runtime_pipeline.jsonl); tf = teacher-forcing forward."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import KOTOBA, synth_audio
from kotoba_whisper_b200 import (LogMelProducer, WhisperB200Config, WhisperB200ForConditionalGeneration,
                                 WhisperFeatureExtractorB200, transcribe_longform)
from kotoba_whisper_b200.pipeline import pipeline
from kotoba_whisper_b200.random_init import random_state_dict
dev = torch.device("cuda", 0)
what = sys.argv[1:] or ["cfg1", "cfg3", "cfg4", "cfg5", "lat", "tf"]
fe = WhisperFeatureExtractorB200(feature_size=128, device=dev)
out = {}

def build(arch, dtype, mb):
    cfg = WhisperB200Config(**arch)
    m = WhisperB200ForConditionalGeneration.from_state_dict(random_state_dict(cfg, 0, dev), cfg, dtype=dtype, max_batch=mb, device=dev)
    torch.cuda.empty_cache()
    return m

def timed(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps, r

if "cfg1" in what:
    m = build(KOTOBA, torch.float32, 4); a = list(synth_audio(4, 1000)); st = {}
    def f():
        x = fe(a, sampling_rate=16000, return_tensors="pt", keep_on_device=True)["input_features"]
        return m.generate(x, language="ja", task="transcribe", return_timestamps=True, max_length=128, stats=st).cpu()
    t, ids = timed(f)
    out["cfg1_fp32_b4_ts"] = {"s_per_batch": t, "rtfx": 120 / t, "passes": st.get("passes"), "ids_shape": list(ids.shape)}
    del m
if "cfg3" in what:
    m = build(dict(KOTOBA, decoder_layers=32), torch.bfloat16, 32); a = list(synth_audio(32, 3000)); st = {}
    def f():
        x = fe(a, sampling_rate=16000, return_tensors="pt", keep_on_device=True)["input_features"]
        return m.generate(x, language="ja", task="transcribe", return_timestamps=True, max_length=128, stats=st).cpu()
    t, ids = timed(f)
    out["cfg3_teacher_bf16_b32_ts"] = {"s_per_batch": t, "rtfx": 32 * 30 / t, "passes": st.get("passes"), "ids_shape": list(ids.shape)}
    # the same labelling loop with two batches in flight (GenerateStream: first pass of batch i under the encoder of i+1)
    stream = m.generate_stream(language="ja", task="transcribe", return_timestamps=True, max_length=128)
    def g():
        x = fe(a, sampling_rate=16000, return_tensors="pt", keep_on_device=True)["input_features"]
        r = stream.submit(x)
        return None if r is None else r.cpu()
    g()
    t, _ = timed(g)
    stream.flush()
    out["cfg3_teacher_bf16_b32_ts_stream"] = {"s_per_batch": t, "rtfx": 32 * 30 / t}
    # coalesced: 4 submitted 32-clip batches run as one 128-row device batch (both greedy passes at 128 rows)
    del stream, m
    torch.cuda.empty_cache()
    m = build(dict(KOTOBA, decoder_layers=32), torch.bfloat16, 128)
    stream = m.generate_stream(coalesce=4, language="ja", task="transcribe", return_timestamps=True, max_length=128)
    for _ in range(8): g()
    t, _ = timed(g, reps=8, warm=0)
    while stream.flush() is not None: pass
    out["cfg3_teacher_bf16_b32_ts_coalesce4"] = {"s_per_batch": t, "rtfx": 32 * 30 / t, "device_batch": 128}
    if "tf" in what:
        x = fe(a, sampling_rate=16000, return_tensors="pt", keep_on_device=True)["input_features"]
        labels = torch.randint(0, 50257, (32, 128), device=dev)
        t, _ = timed(lambda: m(input_features=x, labels=labels).logits)
        enc = m.get_encoder()(x)
        t2, _ = timed(lambda: m(encoder_outputs=enc, labels=labels).logits)
        out["tf_teacher_bf16_b32_T128"] = {"s_with_encoder": t, "s_decoder_only": t2, "rtfx_with_encoder": 32 * 30 / t,
                                           "tokens_per_s_decoder_only": 32 * 128 / t2}
    del m
if "cfg4" in what or "lat" in what:
    m = build(KOTOBA, torch.bfloat16, 64)
    if "cfg4" in what:
        rng = np.random.default_rng(4000); minutes = 60
        audio = (rng.standard_normal(16000 * 60 * minutes) * 0.1).astype(np.float32)
        for mode in (True, False):
            st = {}
            def f():
                return transcribe_longform(m, fe, audio, chunk_length_s=15, batch_size=64, language="ja", task="transcribe",
                                           max_new_tokens=124, device_chunker=mode, stats=st)
            t, toks = timed(f, reps=1, warm=1 if mode else 0)
            out["cfg4_longform_1h_chunk15_b64_" + ("device_chunker" if mode else "host_chunker")] = {
                "s_total": t, "rtfx": 60 * minutes / t, "merged_tokens": len(toks), "windows": st.get("windows"),
                "h2d_bytes": st.get("h2d_bytes")}
        # 128 windows per generate call: the decode-time kernels take up to 128 rows (the latency chain of a decoder
        # position is paid once for twice the rows)
        m128 = build(KOTOBA, torch.bfloat16, 128); st = {}
        def f128():
            return transcribe_longform(m128, fe, audio, chunk_length_s=15, batch_size=128, language="ja", task="transcribe",
                                       max_new_tokens=124, device_chunker=True, stats=st)
        t, toks128 = timed(f128, reps=1, warm=1)
        out["cfg4_longform_1h_chunk15_b128_device_chunker"] = {
            "s_total": t, "rtfx": 60 * minutes / t, "merged_tokens": len(toks128), "windows": st.get("windows"),
            "same_tokens_as_b64": list(toks128) == list(toks)}
        del m128
    if "lat" in what:
        # run_speed_eval.py: batch size 1, chunk_length_s=15, one synthetic clip (rand-0.5)*2*0.007 of `duration` s,
        # language/task in generate_kwargs, wall clock around the whole pipeline call, 1 warm-up + 10 trials
        pipe = pipeline("automatic-speech-recognition", model=m, feature_extractor=fe, chunk_length_s=15)
        ref = {10: 0.0410, 30: 0.1112, 60: 0.2136, 300: 1.0769}  # distil-whisper-bilingual-v1.0 (same arch), runtime_pipeline.jsonl:60,48,36,12
        # Random-init weights never emit eos on noise, so a window decodes to max_length (448); the published numbers were
        # taken with trained weights that stop after a handful of tokens on this input.  Both ends are
        # recorded: max_new_tokens=16 (the published measurement's regime) and the full 444-token decode per window.
        for mnt in (16, None):
            for dur in (10, 30, 60, 300):
                a = ((np.random.rand(int(16000 * dur)) - 0.5) * 2 * 0.007).astype(np.float32)
                gk = {"language": "en", "task": "translate"}
                if mnt: gk["max_new_tokens"] = mnt
                el = []
                for _ in range(11):
                    t0 = time.perf_counter(); pipe(a.copy(), generate_kwargs=dict(gk))
                    el.append(time.perf_counter() - t0)
                el = el[1:]
                out[f"lat_pipeline_b1_{dur}s_" + (f"mnt{mnt}" if mnt else "full448")] = {
                    "mean_s": float(np.mean(el)), "std_s": float(np.std(el)), "rtfx": dur / float(np.mean(el)),
                    "reference_published_s": ref[dur], "speedup_vs_published": ref[dur] / float(np.mean(el))}
    del m
if "cfg5" in what:
    for nm in (80, 128):
        f2 = WhisperFeatureExtractorB200(feature_size=nm, device=dev)
        for B in (1024, 4096, 16384, 65536):
            # resident: slabs of <= 4096 clips in HBM (4096 x 1.92 MB = 7.9 GB in, <= 6.3 GB out), re-used for larger B
            nb = min(B, 4096)
            x = torch.randn(nb, 480000, device=dev) * 0.1
            reps = B // nb
            def run():
                for _ in range(reps): f2.logmel_device(x)
            t, _ = timed(run, reps=2 if B <= 4096 else 1)
            byts = B * (480000 * 4 + nm * 3000 * 4)
            out[f"cfg5_logmel_{nm}_B{B}_resident"] = {"ms": t * 1e3, "clips_per_s": B / t, "GBps": byts / t / 1e9,
                                                       "slab_clips": nb}
            del x
        # producer loop (run_data_filtering's map stage): host clips -> pinned -> H2D -> kernel -> D2H -> host arrays
        prod = LogMelProducer(f2, slab_clips=256)
        host = [np.random.default_rng(i).standard_normal(480000).astype(np.float32) * 0.1 for i in range(256)]
        for B in (1024, 4096):
            batches = [host] * (B // 256)
            def run():
                n = 0
                for _, feats in prod.produce(batches): n += len(feats)
                return n
            t, n = timed(run, reps=1, warm=1)
            out[f"cfg5_logmel_{nm}_B{B}_producer_e2e"] = {"s": t, "clips_per_s": n / t, "h2d_GBps": n * 1.92e-3 / t,
                                                           "d2h_GBps": n * nm * 3000 * 4 / 1e9 / t}
print(json.dumps(out))
