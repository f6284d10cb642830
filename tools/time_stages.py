"""Device time of the pipeline stages (log-mel, encode, greedy pass) at the bench workload, CUDA events, 5 reps."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import KOTOBA, synth_audio  # noqa: E402
from kotoba_whisper_b200 import WhisperB200Config, WhisperB200ForConditionalGeneration, WhisperFeatureExtractorB200  # noqa: E402
from kotoba_whisper_b200.random_init import random_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
cfg = WhisperB200Config(**KOTOBA)
model = WhisperB200ForConditionalGeneration.from_state_dict(random_state_dict(cfg, 0, dev), cfg, dtype=torch.bfloat16,
                                                            max_batch=B, device=dev)
fe = WhisperFeatureExtractorB200(feature_size=128, device=dev)
audio = torch.from_numpy(synth_audio(B, 1000)).to(dev)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for rep in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0 = ev()
    feats = fe.logmel_device(audio)
    e1 = ev()
    model.encode(feats, return_hidden=False)
    e2 = ev()
    c0 = time.perf_counter()
    toks = model._greedy_pass(B, [50258, 50266, 50360, 50364], 128, False)
    c1 = time.perf_counter()
    e3 = ev()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"rep {rep}: logmel {e0.elapsed_time(e1):7.2f} ms  encode {e1.elapsed_time(e2):7.2f} ms  greedy pass {e2.elapsed_time(e3):7.2f} ms "
          f"(host wall of pass {1e3*(c1-c0):7.2f} ms)  total wall {1e3*(t1-t0):7.2f} ms", flush=True)
