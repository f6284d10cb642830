// Device-side pieces of the Whisper logits processors shared by the stand-alone sampling kernel, the vocabulary GEMM's
// fused arg-max epilogue and the combine kernel (HF/generation/logits_process.py:1855-1862, 1898-1902, 1996-2043).
#pragma once
#include "common.cuh"

namespace kw {
namespace sr {

struct Best {
  float v;
  int i;
};
__device__ __forceinline__ Best better(Best a, Best b) {  // larger value wins; ties -> smaller index (torch.argmax)
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ Best warp_best(Best x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = better(x, y);
  }
  return x;
}

// Row state of the timestamp rules, re-derived from the token history (exactly what the HF processor derives from
// input_ids[k, begin_index:]): bit0 at_begin, bit1 last token is a timestamp, bit2 penultimate is a timestamp (or fewer
// than two sampled), bit3 any timestamp so far; bound = first timestamp id still allowed.
__device__ __forceinline__ void row_state(const int* trow, int pos, int begin_index, int tb, int* st, int* bound) {
  const int n = pos + 1 - begin_index;
  const int last_ts = n >= 1 && trow[pos] >= tb;
  const int pen_ts = n < 2 || trow[pos - 1] >= tb;
  int has_ts = 0, ts_last = 0;
  for (int j = pos; j >= begin_index; --j)
    if (trow[j] >= tb) {
      has_ts = 1;
      ts_last = trow[j];
      break;
    }
  *st = (n == 0 ? 1 : 0) | (last_ts << 1) | (pen_ts << 2) | (has_ts << 3);
  *bound = (last_ts && !pen_ts) ? ts_last : ts_last + 1;
}

// Warp-cooperative version (all 32 lanes call it, all get the result): the backwards search for the most recent
// timestamp looks at 32 history positions per step (one ballot) instead of one dependent load per position — the serial
// scan costs ~0.15 us per token of history, 20 us at 124 tokens — and is skipped entirely when the timestamp rules are
// off (only at_begin is read then).
__device__ __forceinline__ void row_state_warp(const int* trow, int pos, int begin_index, int tb, int return_ts, int lane,
                                               int* st, int* bound) {
  const int n = pos + 1 - begin_index;
  if (!return_ts) {
    *st = n == 0 ? 1 : 0;
    *bound = 0;
    return;
  }
  const int last_ts = n >= 1 && trow[pos] >= tb;
  const int pen_ts = n < 2 || trow[pos - 1] >= tb;
  int has_ts = 0, ts_last = 0;
  for (int j0 = pos; j0 >= begin_index && !has_ts; j0 -= 32) {
    const int j = j0 - lane;
    const int tok = j >= begin_index ? trow[j] : 0;
    const unsigned hit = __ballot_sync(0xffffffffu, j >= begin_index && tok >= tb);
    if (hit) {
      has_ts = 1;
      ts_last = __shfl_sync(0xffffffffu, tok, __ffs(hit) - 1);  // lowest lane = most recent position
    }
  }
  *st = (n == 0 ? 1 : 0) | (last_ts << 1) | (pen_ts << 2) | (has_ts << 3);
  *bound = (last_ts && !pen_ts) ? ts_last : ts_last + 1;
}

// f: bit0 = in suppress_tokens, bit1 = in begin_suppress_tokens
__device__ __forceinline__ bool token_masked(const SampleRules& r, int return_ts, int v, unsigned f, int st, int bound) {
  const bool at_begin = st & 1, last_ts = st & 2, pen_ts = st & 4, has_ts = st & 8;
  if (f & 1) return true;
  if (at_begin && (f & 2)) return true;
  if (return_ts) {
    if (v == r.no_ts) return true;
    if (last_ts) {
      if (pen_ts) { if (v >= r.ts_begin) return true; }
      else if (v < r.eos) return true;
    }
    if (has_ts && v >= r.ts_begin && v < bound) return true;
    if (at_begin) {
      if (v < r.ts_begin) return true;
      if (r.max_initial >= 0 && v > r.ts_begin + r.max_initial) return true;
    }
  }
  return false;
}

// token_masked() for TEXT ids (v < ts_begin) factored into a row-independent half (lane bits: what the id is) and an
// id-independent half (column bits: what the row's history says); the id is masked iff the two share a bit.
//   bit0: suppressed always (suppress_tokens, or <|notimestamps|> under the timestamp rules)
//   bit1: in begin_suppress_tokens            x   row is at its first sampled position
//   bit2: id < eos                            x   timestamp rules: last token is a timestamp that opened a segment
//   bit3: any text id                         x   timestamp rules: first sampled position (must be a timestamp)
__device__ __forceinline__ unsigned text_lane_bits(const SampleRules& r, int return_ts, int v, unsigned f) {
  return ((f & 1) || (return_ts && v == r.no_ts) ? 1u : 0u) | ((f & 2) ? 2u : 0u) | (v < r.eos ? 4u : 0u) | 8u;
}
// n = tokens sampled so far in this pass, t0 / t1 = last / second-to-last token of the history
__device__ __forceinline__ int text_col_mask(int return_ts, int n, int t0, int t1, int tb) {
  const bool at_begin = n == 0, last_ts = n >= 1 && t0 >= tb, pen_ts = n < 2 || t1 >= tb;
  return 1 | (at_begin ? 2 : 0) | ((return_ts && last_ts && !pen_ts) ? 4 : 0) | ((return_ts && at_begin) ? 8 : 0);
}

}  // namespace sr
}  // namespace kw
