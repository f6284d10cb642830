"""The fused processors + argmax kernel (kw_sample) vs the oracle's restatement of the three HF logits processors, on
crafted token histories that hit every rule (begin, pairing, monotonicity, max-initial, timestamp-mass, finished rows)."""
import numpy as np
import pytest
import torch

from _gpu_util import build_pair
from _synth import TINY
from kotoba_whisper_b200 import _lib

pytestmark = pytest.mark.gpu
TB, EOS = 50365, 50257


def test_sample_kernel_matches_oracle_rules():
    model, ref = build_pair(TINY, torch.float32, max_batch=4)
    lib = _lib.load()
    V = 51866
    rng = np.random.default_rng(0)
    histories = [[], [TB + 5], [TB + 5, 100], [TB + 5, 100, 200, TB + 40], [TB + 5, 100, TB + 40, TB + 40],
                 [TB, 7, TB + 1500], [TB + 3, 15, 16, 17], [TB + 1499, TB + 1499], [TB + 10, TB + 10, TB + 10 + 1, 42]]
    n_checked = 0
    for rt in (True, False):
        for hist in histories:
            for boost in (None, "ts", "text", "eos"):
                prompt = [50258, 50266, 50360] + ([] if rt else [50364])
                B = 4
                logits = torch.from_numpy(rng.standard_normal((B, V)).astype(np.float32))
                if boost == "ts":
                    logits[:, TB:] += 2.0
                elif boost == "text":
                    logits[:, int(rng.integers(300, 40000))] += 9.0
                elif boost == "eos":
                    logits[:, EOS] += 12.0
                pos = len(prompt) + len(hist) - 1
                ld = pos + 2
                toks = torch.tensor([prompt + hist + [0]] * B, dtype=torch.int32, device="cuda")
                fin = torch.tensor([0, 0, 1, 0], dtype=torch.int32, device="cuda")
                _lib.check(lib.kw_sample(model._handle, logits.cuda().data_ptr(), toks.data_ptr(), ld, B, pos, len(prompt),
                                         int(rt), fin.data_ptr(), torch.cuda.current_stream().cuda_stream), "kw_sample")
                got = toks[:, pos + 1].cpu().tolist()
                sc = ref.process_scores(logits, [list(hist)] * B, rt)
                want = sc.argmax(-1).tolist()
                want[2] = EOS  # finished row emits pad (= eos id)
                assert got == want, (rt, hist, boost, got, want)
                fin_want = [int(w == EOS) for w in want]
                fin_want[2] = 1
                assert fin.cpu().tolist() == fin_want
                n_checked += 1
    assert n_checked == 72


@pytest.mark.parametrize("ts", [False, True])
def test_vocab_projection_with_fused_processors_matches_sample_kernel(ts):
    """EPI_ARGMAX (gemm_tc.cu) + sample_combine vs fp32 logits + sample_kernel: the accumulators are the same bits, so
    with the same token history both must pick the same token at every position — checked on whole greedy passes
    (timestamps on: all pairing / monotonicity / initial-timestamp rules and the log-sum-exp comparison are exercised)."""
    import numpy as np
    from _gpu_util import build_pair
    from _synth import TINY, clips
    from oracle.logmel_ref import logmel_batch_f64
    from kotoba_whisper_b200 import _lib
    lib = _lib.load()
    model, _ = build_pair(TINY, torch.bfloat16, max_batch=8)
    mel = torch.from_numpy(logmel_batch_f64(clips("UGSGSUG", 31), 128)).cuda()
    prompt = [50258, 50266, 50360] + ([] if ts else [50364])
    out = {}
    try:
        for fused in (1, 0):
            lib.kw_set_sample_fused(fused)
            model.encode(mel, return_hidden=False)
            lib.kw_launch_count(1)
            out[fused] = model._greedy_pass(7, prompt, 96, ts)
            out[f"n{fused}"] = lib.kw_launch_count(0)
    finally:
        lib.kw_set_sample_fused(1)
    assert np.array_equal(out[1], out[0]), [(b, int(np.argmax(out[1][b] != out[0][b]))) for b in range(7)
                                            if not np.array_equal(out[1][b], out[0][b])]
    # combine replaces sample, and also does the next position's embedding + first LayerNorm: two launches fewer per
    # sampled position but the last
    assert out["n0"] - 2 * (96 - len(prompt)) <= out["n1"] < out["n0"], (out["n1"], out["n0"])
