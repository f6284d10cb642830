"""Throughput of the other BASELINE.json configs (parity-test cases, not bench lines) on one B200:
   cfg1 fp32 kotoba B=4 timestamps, cfg3 teacher bf16 B=32 timestamps, cfg4 long-form 15 s chunks, cfg5 log-mel sweep."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import KOTOBA, synth_audio
from kotoba_whisper_b200 import (WhisperB200Config, WhisperB200ForConditionalGeneration, WhisperFeatureExtractorB200,
                                 transcribe_longform)
from kotoba_whisper_b200.random_init import random_state_dict
dev = torch.device("cuda", 0)
what = sys.argv[1:] or ["cfg1", "cfg3", "cfg4", "cfg5"]
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
    del m
if "cfg4" in what:
    m = build(KOTOBA, torch.bfloat16, 64)
    rng = np.random.default_rng(4000); minutes = 60
    audio = (rng.standard_normal(16000 * 60 * minutes) * 0.1).astype(np.float32)
    def f():
        return transcribe_longform(m, fe, audio, chunk_length_s=15, batch_size=64, language="ja", task="transcribe", max_new_tokens=124)
    t, toks = timed(f, reps=1, warm=0)
    out["cfg4_longform_1h_chunk15_b64"] = {"s_total": t, "rtfx": 60 * minutes / t, "merged_tokens": len(toks), "chunks": 360}
    del m
if "cfg5" in what:
    for nm in (80, 128):
        f2 = WhisperFeatureExtractorB200(feature_size=nm, device=dev)
        for B in (1024, 4096):
            x = torch.randn(B, 480000, device=dev) * 0.1
            t, _ = timed(lambda: f2.logmel_device(x), reps=3)
            byts = B * (480000 * 4 + nm * 3000 * 4)
            out[f"cfg5_logmel_{nm}_B{B}"] = {"ms": t * 1e3, "clips_per_s": B / t, "GBps": byts / t / 1e9}
            del x
print(json.dumps(out))
