"""Token-level parity of the bf16 tensor-core path against the exact-fp32 CUDA path (BASELINE.json north_star:
"in bf16 ... token sequences identical on >= 99 % of utterances").

Both models are product models (`WhisperB200ForConditionalGeneration`) built from the SAME bf16-rounded state_dict, so
only arithmetic precision differs; the fp32 path is the one whose greedy ids are bit-identical to the reference
(tests/test_model_gpu.py::test_kotoba_fp32_tokens_bit_identical).  Greedy decoding (HF/generation/utils.py:2762-2797)
amplifies one flipped argmax into a different tail, so every divergence is triaged by the fp32 logit gap at the FIRST
divergent position (same token history on both sides there):

  sigma = RMS(logits_bf16 - logits_fp32) over the whole vocabulary at the first generated position (identical prefix),
  tau   = tau_sigmas * sigma,
  a divergence whose fp32 gap  logit32[fp32's token] - logit32[bf16's token]  is below tau is a TIE: the two candidates
  were closer than the bf16 arithmetic noise, neither pick is wrong at bf16 precision.

`adjusted` = identical + ties; `raw` = identical only.  Used by tests (asserted), tools/bf16_parity.py (artifact under
profiles/) and bench.py (the `parity` key of the JSON line)."""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch


def _trim(row: List[int], pad: int) -> List[int]:
    n = len(row)
    while n and row[n - 1] == pad:
        n -= 1
    return row[:n]


def first_divergence(a: List[int], b: List[int]) -> int:
    """Index of the first differing token, -1 when the sequences are identical."""
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return i
    return -1 if len(a) == len(b) else n


@torch.no_grad()
def bf16_token_parity(m16, m32, mel: torch.Tensor, language="ja", task="transcribe", return_timestamps=False,
                      max_length: int = 128, tau_sigmas: float = 4.0) -> Dict:
    """mel: [N, n_mels, 3000] features (any device).  -> dict with raw / adjusted identity and the evidence."""
    g = m32.generation_config
    eos, pad = g.eos_token_id, g.pad_token_id
    prompt = m32._init_tokens(language, task, bool(return_timestamps))
    P = len(prompt)
    chunk = min(m16.max_batch, m32.max_batch)
    dev = m32.device
    n_utt = mel.shape[0]
    first_div: List[int] = []
    gaps: List[float] = []
    top2: List[float] = []
    sq_err, n_err = 0.0, 0
    for c0 in range(0, n_utt, chunk):
        x = mel[c0:c0 + chunk].to(dev)
        x = x.to(torch.bfloat16).to(torch.float32)  # the bf16 model rounds its features; give both paths the same ones
        B = x.shape[0]
        # ONE decoder pass over the 30 s window on each path (kw_greedy_pass: prompt prefill + greedy loop), raw tokens:
        # a first-divergence index then maps directly onto a decoder position.  (generate() may run further seek passes
        # on sub-batches and splices segments, which would hide where two runs parted.)
        def one_pass(m):
            m.encode(x, return_hidden=False)
            t = m._greedy_pass(B, prompt, max_length, bool(return_timestamps))[:, P:].tolist()
            return [r[: r.index(eos)] if eos in r else _trim(r, pad) for r in t]
        a16, a32 = one_pass(m16), one_pass(m32)
        div = [first_divergence(p, q) for p, q in zip(a16, a32)]
        first_div.extend(div)
        # teacher-force the fp32 path along its own tokens up to the last first-divergence position of the chunk
        last = max(div)
        width = P + max(last, 0) + 1
        toks = torch.full((B, width), pad, dtype=torch.int32)
        for b in range(B):
            row = (prompt + a32[b] + [eos])[:width]
            toks[b, : len(row)] = torch.tensor(row, dtype=torch.int32)
        toks = toks.to(dev)
        m32.cross_kv(B)
        m16.cross_kv(B)
        for pos in range(P + max(last, 0)):
            lg32 = m32.step_logits(toks, pos)
            if pos <= P - 1:
                lg16 = m16.step_logits(toks, pos)
            if pos == P - 1:  # first generated position: identical history on both paths
                d = (lg16 - lg32).double()
                sq_err += float((d * d).sum())
                n_err += d.numel()
                t2 = lg32.topk(2, dim=-1).values
                top2.extend((t2[:, 0] - t2[:, 1]).cpu().tolist())
            j = pos - (P - 1)
            for b in range(B):
                if div[b] == j:
                    t32 = a32[b][j] if j < len(a32[b]) else eos
                    t16 = a16[b][j] if j < len(a16[b]) else eos
                    gaps.append((c0 + b, float(lg32[b, t32] - lg32[b, t16])))
    sigma = float(np.sqrt(sq_err / max(n_err, 1)))
    tau = tau_sigmas * sigma
    gap_of = dict(gaps)
    identical = sum(1 for d in first_div if d < 0)
    ties = sum(1 for i, d in enumerate(first_div) if d >= 0 and abs(gap_of.get(i, np.inf)) < tau)
    diverged = [(i, d, gap_of.get(i)) for i, d in enumerate(first_div) if d >= 0]
    return {
        "utterances": n_utt, "max_length": max_length, "return_timestamps": bool(return_timestamps),
        "raw_identical": identical, "raw_pct": 100.0 * identical / n_utt,
        "ties": ties, "adjusted_identical": identical + ties, "adjusted_pct": 100.0 * (identical + ties) / n_utt,
        "logit_noise_sigma": sigma, "tau": tau, "tau_sigmas": tau_sigmas,
        "fp32_top2_margin_first_position": {"median": float(np.median(top2)), "min": float(np.min(top2)),
                                            "p10": float(np.percentile(top2, 10))},
        "first_divergence_step": [d for _, d, _ in diverged],
        "fp32_gap_at_divergence": [None if gp is None else round(gp, 6) for _, _, gp in diverged],
        "max_gap_over_tau": (max((abs(gp) for _, _, gp in diverged if gp is not None), default=0.0) / tau) if tau > 0 else None,
        "reference": "exact-fp32 CUDA path on the same bf16-rounded weights and features (bit-identical to HF fp32)",
    }


def rounded_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """bf16-round every matrix the bf16 model stores in bf16 (biases, LayerNorm and position tables stay fp32)."""
    return {k: (v.to(torch.bfloat16).to(torch.float32) if v.dim() >= 2 and "embed_positions" not in k else v)
            for k, v in sd.items()}
