"""CUDA log-mel kernel (kw_logmel through the C ABI / the FeatureExtractor drop-in) vs the float64 oracle and the HF
goldens.  Tolerance: 1e-5 absolute (BASELINE.json north_star: "Log-mel must match to 1e-5 abs in fp32")."""
import numpy as np
import pytest
import torch

from _synth import clip, clips
from oracle.logmel_ref import frame_attention_mask, logmel_batch_f64, logmel_f64

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def fes():
    from kotoba_whisper_b200 import WhisperFeatureExtractorB200
    return {nm: WhisperFeatureExtractorB200(feature_size=nm, device="cuda:0") for nm in (80, 128)}


@pytest.mark.parametrize("nm", [80, 128])
def test_logmel_matches_oracle_and_goldens(fes, golden, nm):
    cl = clips("UGSS", 500 + nm) + [np.zeros(1000, np.float32)]
    out = fes[nm](cl, sampling_rate=16000, return_tensors="np", return_attention_mask=True)
    x = out["input_features"]
    assert x.shape == (5, nm, 3000) and x.dtype == np.float32
    ref = logmel_batch_f64(cl, nm)
    err = np.abs(x - ref)
    assert err.max() <= TOL, f"max abs err {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"
    g = golden["logmel"]
    assert np.abs(x[:, :, ::37] - g[f"hf_numpy_{nm}"]).max() <= TOL        # HF float64 numpy path
    assert np.abs(x[:, :, ::37] - g[f"hf_torch_{nm}"]).max() <= 5e-5       # HF fp32 torch path: its own error, reported
    assert np.abs(x[:, :, :8] - g[f"hf_numpy_head_{nm}"]).max() <= TOL     # left reflection edge
    assert np.abs(x[:, :, -8:] - g[f"hf_numpy_tail_{nm}"]).max() <= TOL    # right reflection / padding edge
    assert (out["attention_mask"].sum(-1) == g[f"mask_sum_{nm}"]).all()
    assert np.all(x[4] == -1.5)                                            # silence -> exactly the floor


def test_logmel_edge_cases(fes):
    fe = fes[128]
    rng = np.random.default_rng(3)
    # ragged batch incl. 1-sample, exactly-30 s, longer-than-30 s (truncated) and huge-amplitude clips
    cl = [rng.standard_normal(1).astype(np.float32), clip("G", 11), rng.standard_normal(500000).astype(np.float32),
          (rng.standard_normal(123457) * 30000).astype(np.float32), clip("U", 12)[:399], clip("S", 13)]
    x = fe(cl, sampling_rate=16000)["input_features"]
    ref = logmel_batch_f64(cl, 128)
    assert np.abs(x - ref).max() <= TOL
    # single (un-batched) input, list input, float64 input
    one = fe(cl[1].astype(np.float64), sampling_rate=16000)["input_features"]
    assert one.shape == (1, 128, 3000) and np.array_equal(one[0], x[1])    # batched == single, bit for bit
    # non-default padded length that is not a multiple of the hop, "longest" padding
    odd = [clip("G", 21)[:50000], clip("G", 22)[:77777]]
    y = fe(odd, sampling_rate=16000, padding="longest", return_attention_mask=True)
    assert y["input_features"].shape == (2, 128, 77777 // 160)
    assert np.abs(y["input_features"] - logmel_batch_f64(odd, 128, n_samples=77777)).max() <= TOL
    assert np.array_equal(y["attention_mask"], frame_attention_mask([50000, 77777], 77777))
    with pytest.raises(ValueError):
        fe(cl[1], sampling_rate=8000)
    with pytest.raises(ValueError):
        fe(np.zeros((2, 3, 100), np.float32), sampling_rate=16000)


def test_logmel_device_entry_with_lengths(fes):
    fe = fes[80]
    rng = np.random.default_rng(5)
    B, n = 7, 480000
    audio = torch.from_numpy((rng.standard_normal((B, n)) * 0.05).astype(np.float32)).cuda()
    lens = torch.tensor([n, 1, 160, 16000, 479999, 240000, 399], dtype=torch.int32)
    x = fe.logmel_device(audio, lens).cpu().numpy()
    a = audio.cpu().numpy()
    ref = logmel_batch_f64([a[i, : int(lens[i])] for i in range(B)], 80)
    assert np.abs(x - ref).max() <= TOL


def test_logmel_batch64_properties(fes):
    """BASELINE configs[1] batch size: per-clip independence, determinism and the clamp floor at full size."""
    fe = fes[128]
    cl = [clip("UGS"[i % 3], 7000 + i) for i in range(64)]
    x = fe(cl, sampling_rate=16000, return_tensors="pt", keep_on_device=True)["input_features"]
    assert x.shape == (64, 128, 3000)
    again = fe(cl, sampling_rate=16000, return_tensors="pt", keep_on_device=True)["input_features"]
    assert torch.equal(x, again)
    idx = [0, 17, 40, 63]
    sub = fe([cl[i] for i in idx], sampling_rate=16000, return_tensors="pt", keep_on_device=True)["input_features"]
    assert torch.equal(sub, x[idx])
    mx = x.amax(dim=(1, 2), keepdim=True)
    assert bool(((x >= mx - 2.0 - 1e-6)).all())  # (max(v, clipmax-8)+4)/4 spans at most 2.0
    ref = logmel_f64(cl[40], 128)
    assert np.abs(x[40].cpu().numpy() - ref).max() <= TOL


def test_prefetch_matches_blocking_call(fes):
    """`prefetch` (worker-thread staging, side-stream copy + kernel) hands back the same bits as the blocking call,
    also when two prefetches are in flight back to back (the two staging buffers alternate)."""
    fe = fes[128]
    a = [clip("UGS"[i % 3], 9100 + i) for i in range(24)]       # ragged lengths, chunked staging path (B >= 16)
    b = [clip("SGU"[i % 3], 9200 + i) for i in range(5)]        # small-batch path
    want_a = fe(a, sampling_rate=16000, return_tensors="pt", keep_on_device=True)["input_features"].clone()
    want_b = fe(b, sampling_rate=16000, return_tensors="pt", keep_on_device=True)["input_features"].clone()
    pa = fe.prefetch(a, sampling_rate=16000, return_tensors="pt")
    pb = fe.prefetch(b, sampling_rate=16000, return_tensors="pt")
    pa2 = fe.prefetch(a, sampling_rate=16000, return_tensors="pt")
    got_a, got_b, got_a2 = pa.result()["input_features"], pb.result()["input_features"], pa2.result()["input_features"]
    torch.cuda.synchronize()
    assert torch.equal(got_a, want_a) and torch.equal(got_b, want_b) and torch.equal(got_a2, want_a)
