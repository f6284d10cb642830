"""Steady-state step time at B = 64: `model.generate` batch by batch vs GenerateStream (two batches in flight: encoder of
batch i + 1 interleaved with the decoder positions of batch i).  Same work per step; prints ms/step of both and checks
that the token ids agree."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import KOTOBA, synth_audio  # noqa: E402
from kotoba_whisper_b200 import WhisperB200Config, WhisperB200ForConditionalGeneration, WhisperFeatureExtractorB200  # noqa: E402
from kotoba_whisper_b200.random_init import random_state_dict  # noqa: E402

B, N = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda", 0)
cfg = WhisperB200Config(**KOTOBA)
model = WhisperB200ForConditionalGeneration.from_state_dict(random_state_dict(cfg, 0, dev), cfg, dtype=torch.bfloat16,
                                                            max_batch=B, device=dev)
fe = WhisperFeatureExtractorB200(feature_size=128, device=dev)
audio = [torch.from_numpy(synth_audio(B, 1000 + i)).to(dev) for i in range(2)]
kw = dict(language="ja", task="transcribe", return_timestamps=False, max_length=128)


def timed(fn, n):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = [fn(i) for i in range(n)]
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, out


def plain(i):
    return model.generate(fe.logmel_device(audio[i % 2]), **kw)


stream = model.generate_stream(**kw)


def streamed(i):
    return stream.submit(fe.logmel_device(audio[i % 2]))


for _ in range(2):
    timed(plain, 3)
    ms_p, out_p = timed(plain, N)
    timed(streamed, 3)                      # primes the stream (3 = odd: the next submit sees audio[1] ... keep parity)
    ms_s, out_s = timed(streamed, N)
    last = stream.flush()
    print(f"plain {ms_p:.2f} ms/step   stream {ms_s:.2f} ms/step   ({N} steps)", flush=True)
# ids: plain step i used audio[i % 2]; streamed step i returns the batch submitted at step i - 1
ok = all(torch.equal(out_s[i].cpu(), out_p[(i - 1) % 2].cpu()) for i in range(1, N))
print("tokens identical:", ok)
