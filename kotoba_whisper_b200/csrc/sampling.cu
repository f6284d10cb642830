// Logits processors + greedy pick, fused into one pass over the fp32 logit row (one CTA per batch row).
//
// Restates, without host round trips (HF does per-row python loops with .tolist() syncs every step):
//   SuppressTokensAtBeginLogitsProcessor   HF/generation/logits_process.py:1855-1862
//   SuppressTokensLogitsProcessor          :1898-1902
//   WhisperTimeStampLogitsProcessor        :1996-2043  (pairing, monotonicity, max_initial_timestamp, and the
//                                          "logsumexp(timestamps) > max(text)" rule, evaluated in fp32 on the masked row)
//   argmax + finished-row pad + eos bookkeeping of GenerationMixin._sample   HF/generation/utils.py:2793-2805
// The row state (last two sampled tokens' timestamp-ness, last timestamp id, step 0) is re-derived from the token
// history on the device, exactly the quantities the HF processor derives from `input_ids[k, begin_index:]`.
#include <atomic>

#include "common.cuh"

namespace kw {

extern std::atomic<long long> g_launches;


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

constexpr int SM_THREADS = 1024;

__global__ void __launch_bounds__(SM_THREADS)
sample_kernel(const float* __restrict__ logits, const unsigned char* __restrict__ flags, SampleRules r,
              int* __restrict__ tokens, int ld_tokens, int pos, int begin_index, int return_ts,
              int* __restrict__ finished) {
  __shared__ int s_state[5];  // at_begin, last_ts, pen_ts, has_ts, bound
  __shared__ Best s_text[SM_THREADS / 32], s_ts[SM_THREADS / 32];
  __shared__ float s_sum[SM_THREADS / 32];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = logits + (size_t)b * r.vocab;
  int* trow = tokens + (size_t)b * ld_tokens;

  if (tid == 0) {
    const int n = pos + 1 - begin_index;  // tokens sampled so far in this pass
    const int tb = r.ts_begin;
    int last_ts = n >= 1 && trow[pos] >= tb;
    int pen_ts = n < 2 || trow[pos - 1] >= tb;
    int has_ts = 0, ts_last = 0;
    for (int j = pos; j >= begin_index; --j)
      if (trow[j] >= tb) {
        has_ts = 1;
        ts_last = trow[j];
        break;
      }
    s_state[0] = (n == 0);
    s_state[1] = last_ts;
    s_state[2] = pen_ts;
    s_state[3] = has_ts;
    s_state[4] = (last_ts && !pen_ts) ? ts_last : ts_last + 1;
  }
  __syncthreads();
  const bool at_begin = s_state[0], last_ts = s_state[1], pen_ts = s_state[2], has_ts = s_state[3];
  const int bound = s_state[4], tb = r.ts_begin;

  auto masked_f = [&](int v, unsigned char f) -> bool {
    if (f & 1) return true;
    if (at_begin && (f & 2)) return true;
    if (return_ts) {
      if (v == r.no_ts) return true;
      if (last_ts) {
        if (pen_ts) { if (v >= tb) return true; }
        else if (v < r.eos) return true;
      }
      if (has_ts && v >= tb && v < bound) return true;
      if (at_begin) {
        if (v < tb) return true;
        if (r.max_initial >= 0 && v > tb + r.max_initial) return true;
      }
    }
    return false;
  };
  auto masked = [&](int v) -> bool { return masked_f(v, flags[v]); };

  Best bt = {-INFINITY, r.vocab}, bs = {-INFINITY, r.vocab};
  constexpr int U = 8;  // independent loads in flight per thread (the row is 207 KB: latency-, not bandwidth-bound)
  for (int v0 = tid; v0 < r.vocab; v0 += U * SM_THREADS) {
    float x[U];
    unsigned char f[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int v = v0 + u * SM_THREADS;
      x[u] = v < r.vocab ? row[v] : -INFINITY;
      f[u] = v < r.vocab ? flags[v] : (unsigned char)1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int v = v0 + u * SM_THREADS;
      if (v >= r.vocab || masked_f(v, f[u])) continue;
      Best c = {x[u], v};
      if (v < tb) bt = better(bt, c); else bs = better(bs, c);
    }
  }
  bt = warp_best(bt);
  bs = warp_best(bs);
  if (lane == 0) { s_text[warp] = bt; s_ts[warp] = bs; }
  __syncthreads();
  bt = s_text[0];
  bs = s_ts[0];
  for (int w = 1; w < SM_THREADS / 32; ++w) { bt = better(bt, s_text[w]); bs = better(bs, s_ts[w]); }

  int choice;
  if (return_ts) {
    // logsumexp over the unmasked timestamp logits vs the best text logit (log_softmax's shift cancels on both sides)
    float part = 0.0f;
    if (bs.v > -INFINITY)
      for (int v = tb + tid; v < r.vocab; v += SM_THREADS)
        if (!masked(v)) part += expf(row[v] - bs.v);
    part = warp_sum(part);
    if (lane == 0) s_sum[warp] = part;
    __syncthreads();
    float tot = 0.0f;
    for (int w = 0; w < SM_THREADS / 32; ++w) tot += s_sum[w];
    const float lse = (bs.v > -INFINITY) ? bs.v + logf(tot) : -INFINITY;
    if (lse > bt.v) choice = bs.i;
    else choice = (bt.v >= bs.v) ? bt.i : bs.i;
  } else {
    choice = (bt.v >= bs.v) ? bt.i : bs.i;
  }
  if (tid == 0) {
    if (choice >= r.vocab) choice = 0;
    const int fin = finished[b];
    const int next = fin ? r.pad : choice;
    trow[pos + 1] = next;
    if (next == r.eos) finished[b] = 1;
  }
}

int sample_launch(const float* logits, const unsigned char* flags, const SampleRules& r, int* tokens, int ld_tokens,
                  int B, int pos, int begin_index, int return_ts, int* finished, cudaStream_t st) {
  KW_REQUIRE(pos + 1 < ld_tokens && pos + 1 >= begin_index && begin_index >= 1, "sample: pos=%d begin=%d ld=%d", pos,
             begin_index, ld_tokens);
  KW_CUDA_OK(launch_pdl(PDL_SAMPLE, sample_kernel, dim3(B), dim3(SM_THREADS), 0, st, logits, flags, r, tokens, ld_tokens, pos,
                        begin_index, return_ts, finished));
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

}  // namespace kw
