"""The persistent fused decode kernel (csrc/decode_fused.cu) against the kernel-per-op schedule it replaces: same bf16
model, same encoder state, one raw greedy pass each.  The two schedules round identically except for the order of a few
fp32 partial sums (K-group partials, per-warp soft-max partials), so the token streams must be identical or part at a
near-tie of the per-op path's own logits."""
import numpy as np
import pytest
import torch

from _gpu_util import build_pair
from _synth import KOTOBA, TINY, TINY80, clips
from oracle.logmel_ref import logmel_batch_f64

pytestmark = pytest.mark.gpu


def _passes(model, lib, mel, prompt, max_length, ts):
    out = {}
    B = mel.shape[0]
    for impl in (1, 0):
        lib.kw_set_decode_impl(2 if impl else 0)  # 2: the fused kernel or an error, never a silent fallback
        model.encode(mel, return_hidden=False)
        lib.kw_launch_count(1)
        out[impl] = model._greedy_pass(B, prompt, max_length, ts)
        out[f"launches{impl}"] = lib.kw_launch_count(0)
    lib.kw_set_decode_impl(0)
    return out


def _check_near_tie(model, lib, fused, perop, prompt, tol_sigma=6e-3):
    """rows that part: gap of the per-op path's logits between the two picks at the first divergent position"""
    B, P = fused.shape[0], len(prompt)
    lib.kw_set_decode_impl(0)
    try:
        model.cross_kv(B)
        worst = 0.0
        for b in range(B):
            a, c = fused[b].tolist(), perop[b].tolist()
            j = next((i for i in range(P, len(a)) if a[i] != c[i]), -1)
            if j < 0:
                continue
            hist = torch.tensor([c[:j]] * B, dtype=torch.int32, device="cuda")
            for pos in range(j):
                lg = model.step_logits(hist, pos)
            gap = abs(float(lg[b, c[j]] - lg[b, a[j]]))
            scale = float(lg[b].abs().max())
            worst = max(worst, gap / scale)
            assert gap <= tol_sigma * scale, (b, j, a[j], c[j], gap, scale)
        return worst
    finally:
        lib.kw_set_decode_impl(0)


@pytest.mark.parametrize("arch", [TINY, TINY80])
@pytest.mark.parametrize("ts", [False, True])
def test_fused_pass_matches_per_op_schedule_tiny(arch, ts):
    from kotoba_whisper_b200 import _lib
    lib = _lib.load()
    model, _ = build_pair(arch, torch.bfloat16, max_batch=8)
    mel = torch.from_numpy(logmel_batch_f64(clips("UGSGSU", 70), arch["num_mel_bins"])).cuda()
    prompt = [50258, 50266, 50360] + ([] if ts else [50364])
    for B in (6, 1):
        r = _passes(model, lib, mel[:B], prompt, 64, ts)
        assert r["launches1"] < 10 and r["launches0"] > 500, (r["launches1"], r["launches0"])
        same = sum(int(np.array_equal(r[1][b], r[0][b])) for b in range(B))
        model.encode(mel[:B], return_hidden=False)
        _check_near_tie(model, lib, r[1], r[0], prompt)
        assert same >= (B + 1) // 2, (same, B)


def test_fused_pass_generate_surface_tiny():
    """generate() end to end on the fused kernel: timestamps, second seek pass on a sub-batch, early exit on eos."""
    from kotoba_whisper_b200 import _lib
    lib = _lib.load()
    model, _ = build_pair(TINY, torch.bfloat16, max_batch=4)
    mel = torch.from_numpy(logmel_batch_f64(clips("UGS", 7), 128)).cuda()
    res = {}
    for impl in (1, 0):
        lib.kw_set_decode_impl(2 if impl else 0)
        st = {}
        res[impl] = (model.generate(mel, language="ja", task="transcribe", return_timestamps=True, max_length=48, stats=st), st)
    lib.kw_set_decode_impl(0)
    a, b = res[1][0], res[0][0]
    assert a.dtype == torch.long and a.shape[0] == 3
    assert res[1][1]["passes"] >= 1
    same = sum(int(a.shape == b.shape and torch.equal(a[i], b[i])) for i in range(3))
    assert same >= 2, (a, b)


def test_fused_pass_matches_per_op_schedule_kotoba():
    """Full-size decoder (d = 1280, 20 heads, ffn 5120, vocabulary 51866) at batch 64: projections split over 148 CTAs with
    4 K-groups, 1280 (batch, head) pairs of cross-attention, 433 vocabulary tiles."""
    from kotoba_whisper_b200 import _lib
    lib = _lib.load()
    arch = dict(KOTOBA, encoder_layers=2)  # the decoder is what is under test; a 2-layer encoder keeps the test short
    model, _ = build_pair(arch, torch.bfloat16, max_batch=64)
    base = torch.from_numpy(logmel_batch_f64(clips("UGSG", 1000), 128))
    mel = torch.cat([base * (1.0 + 0.01 * i) for i in range(16)]).cuda()
    for ts, prompt in ((False, [50258, 50266, 50360, 50364]), (True, [50258, 50266, 50360])):
        r = _passes(model, lib, mel, prompt, 48, ts)
        assert r["launches1"] < 10 and r["launches0"] > 500, (r["launches1"], r["launches0"])
        same = sum(int(np.array_equal(r[1][b], r[0][b])) for b in range(64))
        model.encode(mel, return_hidden=False)
        worst = _check_near_tie(model, lib, r[1], r[0], prompt)
        assert same >= 48, (ts, same, worst)
