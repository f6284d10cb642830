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
#include "ln_row.cuh"
#include "sample_rules.cuh"

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
constexpr int SM_SPLIT = 4;  // CTAs (one thread-block cluster) per batch row: 64 rows alone would leave 84 SMs idle

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_cluster_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(SM_THREADS)
sample_kernel(const float* __restrict__ logits, const unsigned char* __restrict__ flags, SampleRules r,
              int* __restrict__ tokens, int ld_tokens, int pos, int begin_index, int return_ts,
              int* __restrict__ finished) {
  __shared__ int s_state[5];  // at_begin, last_ts, pen_ts, has_ts, bound
  __shared__ Best s_text[SM_THREADS / 32], s_ts[SM_THREADS / 32];
  __shared__ float s_sum[SM_THREADS / 32];
  __shared__ Best s_part_text[SM_SPLIT], s_part_ts[SM_SPLIT];  // per-CTA partials, gathered in the cluster's CTA 0
  __shared__ float s_part_sum[SM_SPLIT];
  pdl_trigger();
  pdl_wait();
  const int crank = (int)cluster_rank();
  const int b = blockIdx.x / SM_SPLIT, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // this CTA's slice of the vocabulary
  const int slice = (r.vocab + SM_SPLIT - 1) / SM_SPLIT, v_lo = crank * slice, v_hi = min(r.vocab, v_lo + slice);
  const float* row = logits + (size_t)b * r.vocab;
  int* trow = tokens + (size_t)b * ld_tokens;

  if (tid < 32) {  // warp 0: row state, 32 history positions per step (sample_rules.cuh)
    int st, bd;
    sr::row_state_warp(trow, pos, begin_index, r.ts_begin, return_ts, tid, &st, &bd);
    if (tid == 0) {
      s_state[0] = st & 1;
      s_state[1] = (st >> 1) & 1;
      s_state[2] = (st >> 2) & 1;
      s_state[3] = (st >> 3) & 1;
      s_state[4] = bd;
    }
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
  for (int v0 = v_lo + tid; v0 < v_hi; v0 += U * SM_THREADS) {
    float x[U];
    unsigned char f[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int v = v0 + u * SM_THREADS;
      x[u] = v < v_hi ? row[v] : -INFINITY;
      f[u] = v < v_hi ? flags[v] : (unsigned char)1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int v = v0 + u * SM_THREADS;
      if (v >= v_hi || masked_f(v, f[u])) continue;
      Best c = {x[u], v};
      if (v < tb) bt = better(bt, c); else bs = better(bs, c);
    }
  }
  bt = warp_best(bt);
  bs = warp_best(bs);
  if (lane == 0) { s_text[warp] = bt; s_ts[warp] = bs; }
  __syncthreads();
  // CTA partials -> CTA 0 of the cluster (distributed shared memory), then every CTA reads the row-wide result back
  if (tid == 0) {
    bt = s_text[0];
    bs = s_ts[0];
    for (int w = 1; w < SM_THREADS / 32; ++w) { bt = better(bt, s_text[w]); bs = better(bs, s_ts[w]); }
    const uint32_t pt = map_to_rank(smem_addr(&s_part_text[crank]), 0), ps = map_to_rank(smem_addr(&s_part_ts[crank]), 0);
    st_cluster_u32(pt, __float_as_uint(bt.v));
    st_cluster_u32(pt + 4, (uint32_t)bt.i);
    st_cluster_u32(ps, __float_as_uint(bs.v));
    st_cluster_u32(ps + 4, (uint32_t)bs.i);
  }
  cluster_barrier();
  if (!return_ts && crank != 0) return;  // only the timestamp rule needs a second, row-wide pass
  {
    const uint32_t pt = map_to_rank(smem_addr(&s_part_text[0]), 0), ps = map_to_rank(smem_addr(&s_part_ts[0]), 0);
    bt.v = __uint_as_float(ld_cluster_u32(pt)); bt.i = (int)ld_cluster_u32(pt + 4);
    bs.v = __uint_as_float(ld_cluster_u32(ps)); bs.i = (int)ld_cluster_u32(ps + 4);
    for (int c = 1; c < SM_SPLIT; ++c) {  // rank order: ties keep the smaller index, as torch.argmax
      Best t2, s2;
      t2.v = __uint_as_float(ld_cluster_u32(pt + 8 * c)); t2.i = (int)ld_cluster_u32(pt + 8 * c + 4);
      s2.v = __uint_as_float(ld_cluster_u32(ps + 8 * c)); s2.i = (int)ld_cluster_u32(ps + 8 * c + 4);
      bt = better(bt, t2);
      bs = better(bs, s2);
    }
  }

  int choice;
  if (return_ts) {
    // logsumexp over the unmasked timestamp logits vs the best text logit (log_softmax's shift cancels on both sides)
    float part = 0.0f;
    if (bs.v > -INFINITY)
      for (int v = max(tb, v_lo) + tid; v < v_hi; v += SM_THREADS)
        if (!masked(v)) part += expf(row[v] - bs.v);
    part = warp_sum(part);
    if (lane == 0) s_sum[warp] = part;
    __syncthreads();
    if (tid == 0) {
      float ctot = 0.0f;
      for (int w = 0; w < SM_THREADS / 32; ++w) ctot += s_sum[w];
      st_cluster_u32(map_to_rank(smem_addr(&s_part_sum[crank]), 0), __float_as_uint(ctot));
    }
    cluster_barrier();
    if (crank != 0) return;
    float tot = 0.0f;
    for (int c = 0; c < SM_SPLIT; ++c) tot += s_part_sum[c];
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

// Second half of the fused vocabulary projection (EPI_ARGMAX, gemm_tc.cu): one CTA per batch row folds the text-slice
// partials in a fixed order, runs the logits processors over the raw tail (timestamp ids and the specials next to them,
// ~1.5 k logits), applies the "timestamps outweigh text" rule and writes the next token / finished flag
// (GenerationMixin._sample tail).  With en.on it also prepares the next decoder position: x = E[next] + P[pos + 1] and
// LayerNorm_1 of decoder layer 0 (modeling_whisper.py:738, 755-763, 468) — bit-identical to embed_kernel + layernorm_kernel.
constexpr int SC_THREADS = 256;
__global__ void __launch_bounds__(SC_THREADS)
sample_combine_kernel(const SampleFuse sf, int* __restrict__ tokens, int* __restrict__ finished, const EmbedNext en) {
  __shared__ Best s_t[SC_THREADS / 32], s_s[SC_THREADS / 32];
  __shared__ float s_sum[SC_THREADS / 32];
  __shared__ int s_state[2], s_next;
  pdl_trigger();
  pdl_wait();
  const SampleRules& r = sf.rules;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, tb = r.ts_begin;
  int* trow = tokens + (size_t)b * sf.ld_tokens;
  if (warp == 0) {
    int st, bd;
    sr::row_state_warp(trow, sf.pos, sf.begin_index, tb, sf.return_ts, lane, &st, &bd);
    if (lane == 0) { s_state[0] = st; s_state[1] = bd; }
  }
  Best bt = {-INFINITY, r.vocab}, bs = {-INFINITY, r.vocab};
  const float2* vp = sf.vpart + (size_t)b * sf.n_part;
  for (int i = tid; i < sf.n_part; i += SC_THREADS) {
    const float2 o = vp[i];
    Best t = {o.x, __float_as_int(o.y)};
    bt = better(bt, t);
  }
  __syncthreads();
  const int st = s_state[0], bound = s_state[1];
  const float* tl = sf.tail + (size_t)b * sf.tail_ld;
  for (int v = sf.tail0 + tid; v < r.vocab; v += SC_THREADS) {
    if (sr::token_masked(r, sf.return_ts, v, sf.flags[v], st, bound)) continue;
    Best c = {tl[v - sf.tail0], v};
    if (v < tb) bt = better(bt, c); else bs = better(bs, c);
  }
  bt = warp_best(bt);
  bs = warp_best(bs);
  if (lane == 0) { s_t[warp] = bt; s_s[warp] = bs; }
  __syncthreads();
  bt = s_t[0];
  bs = s_s[0];
  for (int w = 1; w < SC_THREADS / 32; ++w) { bt = better(bt, s_t[w]); bs = better(bs, s_s[w]); }
  int choice;
  if (sf.return_ts) {
    // logsumexp over the unmasked timestamp logits vs the best text logit (log_softmax's shift cancels on both sides)
    float part = 0.0f;
    if (bs.v > -INFINITY)
      for (int v = max(tb, sf.tail0) + tid; v < r.vocab; v += SC_THREADS)
        if (!sr::token_masked(r, sf.return_ts, v, sf.flags[v], st, bound)) part += expf(tl[v - sf.tail0] - bs.v);
    part = warp_sum(part);
    if (lane == 0) s_sum[warp] = part;
    __syncthreads();
    float tot = 0.0f;
    for (int w = 0; w < SC_THREADS / 32; ++w) tot += s_sum[w];
    const float lse = (bs.v > -INFINITY) ? bs.v + logf(tot) : -INFINITY;
    choice = (lse > bt.v) ? bs.i : ((bt.v >= bs.v) ? bt.i : bs.i);
  } else {
    choice = (bt.v >= bs.v) ? bt.i : bs.i;
  }
  if (tid == 0) {
    if (choice >= r.vocab) choice = 0;
    const int fin = finished[b];
    const int next = fin ? r.pad : choice;
    trow[sf.pos + 1] = next;
    if (next == r.eos) finished[b] = 1;
    s_next = next;
  }
  if (!en.on) return;
  __syncthreads();
  if (warp != 0) return;
  const int d = en.d, tok = min(max(s_next, 0), r.vocab - 1);
  const bf16* e = reinterpret_cast<const bf16*>(en.E) + (size_t)tok * d;
  const float* pp = en.P + (size_t)(sf.pos + 1) * d;
  constexpr int NV = 16;  // d <= 2048 (checked at launch)
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < d) {
      const float4 a = ld4(e + c), q = *reinterpret_cast<const float4*>(pp + c);
      v[i] = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
      *reinterpret_cast<float4*>(en.x + (size_t)b * d + c) = v[i];
    }
  }
  ln_row<bf16, NV, false>(v, en.ln_w, en.ln_b, reinterpret_cast<bf16*>(en.da) + (size_t)b * d, d, lane);
}

int sample_combine_launch(const SampleFuse& sf, int* tokens, int B, int* finished, const EmbedNext& en, cudaStream_t st) {
  KW_REQUIRE(!en.on || (en.d % 4 == 0 && en.d <= 2048), "sample_combine: d=%d unsupported", en.d);
  KW_CUDA_OK(launch_pdl(PDL_SAMPLE, sample_combine_kernel, dim3(B), dim3(SC_THREADS), 0, st, sf, tokens, finished, en));
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

int sample_launch(const float* logits, const unsigned char* flags, const SampleRules& r, int* tokens, int ld_tokens,
                  int B, int pos, int begin_index, int return_ts, int* finished, cudaStream_t st) {
  KW_REQUIRE(pos + 1 < ld_tokens && pos + 1 >= begin_index && begin_index >= 1, "sample: pos=%d begin=%d ld=%d", pos,
             begin_index, ld_tokens);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(B * SM_SPLIT);
  cfg.blockDim = dim3(SM_THREADS);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = SM_SPLIT;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  KW_CUDA_OK(cudaLaunchKernelEx(&cfg, sample_kernel, logits, flags, r, tokens, ld_tokens, pos, begin_index, return_ts,
                                finished));
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

}  // namespace kw
