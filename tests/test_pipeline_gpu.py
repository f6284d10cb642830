"""Long-form chunked transcription (BASELINE configs[3] shape at test scale) and the configs[4] log-mel sweep properties."""
import numpy as np
import pytest
import torch

from _gpu_util import build_pair
from _synth import TINY, clip
from oracle.chunking_ref import transcribe_longform_ref

pytestmark = pytest.mark.gpu


def test_longform_chunked_pipeline_matches_oracle():
    from kotoba_whisper_b200 import WhisperFeatureExtractorB200, transcribe_longform
    model, ref = build_pair(TINY, torch.float32, max_batch=4)
    fe = WhisperFeatureExtractorB200(feature_size=128, device="cuda:0")
    rng = np.random.default_rng(11)
    audio = (rng.standard_normal(16000 * 64 + 777) * 0.1).astype(np.float32)   # 6 windows of 15 s, last one short
    merged, per_chunk, strides = transcribe_longform(model, fe, audio, chunk_length_s=15, batch_size=3, language="ja",
                                                     task="transcribe", max_new_tokens=24, return_chunk_tokens=True)
    want, want_chunks = transcribe_longform_ref(ref, audio, 128, language="ja", task="transcribe", max_length=28)
    assert len(per_chunk) == len(want_chunks) == 6
    assert strides[0] == (240000, 0, 40000) and strides[-1][2] == 0
    assert per_chunk == want_chunks
    assert merged == want


@pytest.mark.parametrize("nm", [80, 128])
def test_logmel_sweep_properties(nm):
    """1k clips resident in HBM: every clip equals its stand-alone result (spot-checked), values bounded by the clamp."""
    from kotoba_whisper_b200 import WhisperFeatureExtractorB200
    from oracle.logmel_ref import logmel_f64
    fe = WhisperFeatureExtractorB200(feature_size=nm, device="cuda:0")
    B = 1024
    g = torch.Generator(device="cuda").manual_seed(nm)
    audio = torch.randn(B, 480000, generator=g, device="cuda") * 0.05
    lens = torch.randint(400, 480001, (B,), generator=g, device="cuda", dtype=torch.int32)
    out = fe.logmel_device(audio, lens)
    assert out.shape == (B, nm, 3000)
    assert torch.isfinite(out).all()
    mx = out.amax(dim=(1, 2))
    mn = out.amin(dim=(1, 2))
    assert bool((mx - mn <= 2.0 + 1e-6).all())
    for i in (0, 511, 1023):
        a = audio[i, : int(lens[i])].cpu().numpy()
        assert np.abs(out[i].cpu().numpy() - logmel_f64(a, nm)).max() <= 1e-5
