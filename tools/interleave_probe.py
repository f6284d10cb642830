"""Does the SM clock the power governor grants depend on how the encoder-like (power-capped GEMM / attention) and the
decode-like (latency-bound, low power) work are ordered in time?  Same kernels, same count, two orders:
  A  phases:       16 x [2 encoder layers + stem]  then  16 x [greedy pass of 11 positions]
  B  interleaved:  16 x ([2 encoder layers + stem], [greedy pass of 11 positions])
If B is faster than A the step could be sped up by interleaving batch i's decode with batch i+1's encoder."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import KOTOBA, synth_audio  # noqa: E402
from kotoba_whisper_b200 import WhisperB200Config, WhisperB200ForConditionalGeneration, WhisperFeatureExtractorB200, _lib  # noqa: E402
from kotoba_whisper_b200.random_init import random_state_dict  # noqa: E402

B, SL = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
lib = _lib.load()


def build(enc_layers):
    cfg = WhisperB200Config(**dict(KOTOBA, encoder_layers=enc_layers))
    return WhisperB200ForConditionalGeneration.from_state_dict(random_state_dict(cfg, 0, dev), cfg, dtype=torch.bfloat16,
                                                               max_batch=B, device=dev)


full, enc2 = build(2), build(2)   # `full` only decodes here (its 2-layer encoder output feeds the cross K/V)
fe = WhisperFeatureExtractorB200(feature_size=128, device=dev)
feats = fe.logmel_device(torch.from_numpy(synth_audio(B, 1000)).to(dev))
full.encode(feats, return_hidden=False)
full._greedy_pass(B, [50258, 50266, 50360, 50364], 12, False)
tokens = torch.empty((B, 12), dtype=torch.int32, device=dev)
pr = (C.c_int32 * 4)(50258, 50266, 50360, 50364)


def dec():
    _lib.check(lib.kw_greedy_pass(full._handle, B, pr, 4, 12, 0, 0, tokens.data_ptr(), full._stream()), "pass")


def enc():
    enc2.encode(feats, return_hidden=False)


def timed(fn, reps=6):
    ts = []
    for r in range(reps + 2):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if r >= 2:
            ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts), min(ts)


def phases():
    for _ in range(SL): enc()
    for _ in range(SL): dec()


def interleaved():
    for _ in range(SL):
        enc(); dec()


def only_enc():
    for _ in range(SL): enc()


def only_dec():
    for _ in range(SL): dec()


for name, fn in (("encoder slices alone", only_enc), ("decode slices alone", only_dec), ("A phases", phases),
                 ("B interleaved", interleaved), ("A phases", phases), ("B interleaved", interleaved)):
    m, mn = timed(fn)
    print(f"{name:24s} mean {m:8.2f} ms  min {mn:8.2f} ms", flush=True)
