"""One un-warmed pass of the bench workload (log-mel -> encode -> greedy generate) for ncu launch lists / captures.
    python tools/prof_step.py [--batch 64] [--max-length 128] [--dtype bf16]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import KOTOBA, SR, synth_audio  # noqa: E402
from kotoba_whisper_b200 import WhisperB200Config, WhisperB200ForConditionalGeneration, WhisperFeatureExtractorB200  # noqa: E402
from kotoba_whisper_b200.random_init import random_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--max-length", type=int, default=128)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--stream", action="store_true", help="GenerateStream: the profiled region is one steady-state stream step")
ap.add_argument("--enc-layers", type=int, default=0, help="shrink the encoder (same kernel shapes) for --set full captures")
ap.add_argument("--coalesce", type=int, default=1, help="with --stream: submitted batches per device batch (bench.py default: 2); "
                "the profiled region is then `coalesce` steady-state steps = one device-batch launch")
a = ap.parse_args()
co = a.coalesce if a.stream else 1
dev = torch.device("cuda", 0)
cfg = WhisperB200Config(**dict(KOTOBA, encoder_layers=a.enc_layers or KOTOBA["encoder_layers"]))
model = WhisperB200ForConditionalGeneration.from_state_dict(
    random_state_dict(cfg, 0, dev), cfg, dtype=torch.bfloat16 if a.dtype == "bf16" else torch.float32,
    max_batch=a.batch * co, device=dev)
fe = WhisperFeatureExtractorB200(feature_size=128, device=dev)
audio = torch.from_numpy(synth_audio(a.batch, 1000)).to(dev)
kw = dict(language="ja", task="transcribe", return_timestamps=False, max_length=a.max_length)
stream = model.generate_stream(coalesce=co, **kw) if a.stream else None
if stream:
    for _ in range(co):
        stream.submit(fe.logmel_device(audio))   # primes the stream: the profiled steps have a decode to interleave
torch.cuda.synchronize()
torch.cuda.profiler.start()   # ncu --profile-from-start off: skip the random-init kernels
for _ in range(a.steps * co):
    feats = fe.logmel_device(audio)
    ids = stream.submit(feats) if stream else model.generate(feats, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
if stream:
    while stream.flush() is not None:
        pass
print("ok", None if ids is None else tuple(ids.shape))
