// Persistent fused greedy-decode kernel (bf16 models): ONE cooperative launch runs every decoder position of a greedy
// pass — embedding, LayerNorms, QKV / out / cross-q / cross-out / fc1 / fc2 projections on tcgen05, self-attention with
// in-place KV append, cross-attention over the cached encoder K/V, the tied vocabulary projection with the three Whisper
// logits processors and the argmax fused into its epilogue, eos / pad bookkeeping and the all-rows-finished exit —
// replacing the ~27 dependent launches per position of the kernel-per-op schedule (api.cu decode_hidden).
// Restates HF/models/whisper/modeling_whisper.py:449-506 (decoder layer), :734-796 (decoder), :1081 (proj_out),
// HF/generation/logits_process.py:1855-1862, 1898-1902, 1996-2043 and HF/generation/utils.py:2762-2805 (_sample tail).
//
// One CTA per SM (148), 20 warps:
//   warps 0-15  compute: GEMM epilogues (tcgen05.ld -> smem transpose -> global), row-owner phases (residual + LayerNorm,
//               embedding, token pick), both attention phases, vocabulary epilogue (processors + per-warp arg-max partials)
//   warp 16     weight producer: TMA loads of this CTA's weight slices for EVERY projection of EVERY position into a
//               6 x 16 KB shared-memory ring.  Weights depend on nothing, so this warp never takes part in a barrier: it
//               runs ahead of the compute phases by the depth of the ring (the next projections' weights are already in
//               shared memory when their phase starts; during the long cross-attention phase the ring is simply full)
//   warp 17     cross-K/V producer: TMA loads (125 key rows x 128 B per tile) of the encoder K / V rows of this CTA's
//               (batch, head) pairs into a 4 x 16 KB ring, also free-running: the first tiles of a layer's cross-attention
//               are in flight while the preceding phases still run, and HBM never waits for the softmax of a pair
//   warp 18     activation producer: after each phase barrier, TMA loads of the [batch, 64] activation k-blocks
//   warp 19     MMA issuer: tcgen05.mma M=128 (weight rows) x N=64 (batch) x K=16, accumulators in TMEM (2 x 64 columns)
// Phases are separated by a grid-wide barrier (red.release / ld.acquire on one global counter, ~1.3 us on 148 CTAs).
// Work split: every projection is cut over all CTAs — output features for QKV / fc1 / vocabulary, output features x 4
// K-groups for the d-wide outputs (out, cross-q, cross-out, fc2), whose raw fp32 partial sums are reduced in fixed order
// by their consumer (the row-owner LayerNorm phase adds bias + residual, the cross-attention phase adds the q bias) —
// deterministic, no atomics.  CTA b < B owns batch row b for the row-wise phases.
#include <atomic>
#include <vector>

#include "model.cuh"
#include "tc_common.cuh"

namespace kw {

extern std::atomic<long long> g_launches;

namespace fd {
using namespace tc;

constexpr int BK = 64, NB = 64, UMMA_K = 16;
constexpr int N_CWARPS = 16, N_CTHREADS = N_CWARPS * 32;
constexpr int W_WARP = 16, KV_WARP = 17, A_WARP = 18, MMA_WARP = 19;
constexpr int THREADS = 20 * 32;
constexpr int SYNC_THREADS = N_CTHREADS + 64;  // compute + activation + MMA warps take part in the phase barriers
constexpr int W_SLOT = 16384, NSW = 6, KV_SLOT = 16384, NSKV = 4, A_SLOT = NB * BK * 2, NSA = 4, SCRATCH = 16384;
constexpr int OFF_W = 0, OFF_KV = OFF_W + NSW * W_SLOT, OFF_A = OFF_KV + NSKV * KV_SLOT, OFF_SCR = OFF_A + NSA * A_SLOT;
constexpr int OFF_BAR = OFF_SCR + SCRATCH;
constexpr int N_BARS = 2 * NSW + 2 * NSKV + 2 * NSA + 4;
constexpr int OFF_MISC = OFF_BAR + 8 * N_BARS;  // tmem slot, stop flag, all-finished flag
constexpr size_t SMEM_BYTES = 1024 + OFF_MISC + 64;
constexpr int TMEM_COLS = 128;
constexpr uint32_t IDESC = make_idesc(128, NB, 0, 0);
constexpr int MAX_SPLIT = 8, HD = 64, MAX_T = 512, MAX_S = 1536;
constexpr int VP_WORDS = 5;  // vocabulary partial: best text (value, id), best timestamp (value, id), sum exp(ts - best ts)

struct GemmCfg {
  int N, K;      // output features, reduction length
  int R;         // weight rows (output features) per CTA tile
  int rpad;      // R rounded up to the 8-row swizzle atom: row pitch of a k-block inside a ring slot
  int kbps;      // k-blocks packed into one 16 KB ring slot
  int S;         // K-groups (split-K factor); partial sums are reduced by the consumer
  int n_groups;  // feature groups; CTA c -> group c % n_groups, K-group c / n_groups
  int nkb;       // k-blocks per CTA tile
  int tiles;     // vocabulary only: number of R-row tiles (walked c, c + G, ...)
};

struct LayerP {
  const float *ln1_w, *ln1_b, *lnx_w, *lnx_b, *ln3_w, *ln3_b;
  const float *bqkv, *bo, *bq_x, *bo_x, *b1, *b2;
  bf16 *self_k, *self_v;  // [B_max][H][MT][64]
};

enum { M_WQKV = 0, M_WO, M_WQX, M_WOX, M_W1, M_W2, M_XKV, MAPS_PER_LAYER };
enum { M_VOCAB = 0, M_DA, M_DATTN, M_DH, MAPS_GLOBAL };

struct Params {
  const CUtensorMap* maps;  // device array: [L][MAPS_PER_LAYER] then [MAPS_GLOBAL]
  const LayerP* layers;     // device array [L]
  int L, B, d, H, F, V, S, MT, G;
  GemmCfg g_qkv, g_dd, g_fc1, g_fc2, g_voc;
  float *x, *partqkv, *part, *partq, *vpart;  // partqkv [S][64][3d], part / partq [S][64][d]: K-group partial sums
  bf16 *da, *dattn, *dh;
  const bf16* tok_embed;
  const float *dec_pos, *lnf_w, *lnf_b;
  const unsigned char* flags;
  SampleRules rules;
  int* tokens;
  int ld_tokens, n_prompt, max_length, return_ts;
  int* finished;
  unsigned* bar_counter;
  int* steps_out;
  int TR, n_kv_tiles;  // cross K/V tile rows, tiles per (batch, head)
  int kv_sw128;        // cross K/V tiles land 128B-swizzled (16-byte chunk c of row r sits at chunk c ^ (r & 7))
  int dbg_skip;        // bring-up: bit0 skip cross-attention (+ its producer), bit1 skip the vocabulary phase
  unsigned* dbg;       // bring-up breadcrumbs in mapped host memory (16 words per CTA), or nullptr
};

__device__ unsigned* g_dbg = nullptr;  // set by thread 0 of every CTA from Params (read by the watchdogs)

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

__device__ __forceinline__ float gelu_as(float x) {  // GELU(erf), Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7), as gemm_tc.cu
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
  return 0.5f * x * (1.0f + copysignf(fmaf(-poly, e, 1.0f), x));
}

__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// mbarrier wait with a watchdog that names the waiting site (one line per warp) and traps instead of hanging the GPU
__device__ __forceinline__ void fd_wait(uint32_t bar, uint32_t parity, int site) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 3000000000LL) {
      if (g_dbg && ((threadIdx.x & 31) == 0)) {
        g_dbg[blockIdx.x * 16 + 4 + ((threadIdx.x >> 5) >= N_CWARPS ? (threadIdx.x >> 5) - N_CWARPS : 4)] =
            0x80000000u | (unsigned)site | (parity << 8) | ((threadIdx.x >> 5) << 16);
        __threadfence_system();
      }
      if ((threadIdx.x & 31) == 0 || threadIdx.x >= N_CTHREADS)
        printf("kwb200 decode_fused: wait site %d timed out (block %d warp %d parity %u)\n", site, blockIdx.x,
               threadIdx.x >> 5, parity);
      __trap();
    }
  }
}

// mbarrier wait that gives up when the CTA's stop flag is raised (free-running producers); true = barrier completed
__device__ __forceinline__ bool mbar_wait_or_stop(uint32_t bar, uint32_t parity, volatile int* stop) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (*stop) return false;
    if (clock64() - t0 > 6000000000LL) {
      if (g_dbg) {
        g_dbg[blockIdx.x * 16 + 4 + (threadIdx.x >> 5) - N_CWARPS] = 0xA0000000u | (parity << 8);
        __threadfence_system();
      }
      printf("kwb200 decode_fused: producer wait timed out (block %d warp %d)\n", blockIdx.x, threadIdx.x >> 5);
      __trap();
    }
  }
  return true;
}

__device__ __forceinline__ void unpack8(const uint4& r, float* f) {
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h2[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

struct Plan {
  int active, n0, kb0, kg;
};
__device__ __forceinline__ Plan plan_of(const GemmCfg& g, int cta) {
  Plan pl;
  const int ng = cta % g.n_groups;
  pl.kg = cta / g.n_groups;
  pl.n0 = ng * g.R;
  pl.kb0 = pl.kg * g.nkb;
  pl.active = pl.kg < g.S && pl.n0 < g.N;
  return pl;
}

// Everything one CTA needs to address its rings and barriers.
struct Ctx {
  uint32_t base;       // 1024-aligned shared::cta address of the ring area
  uint8_t* gen;        // generic pointer to the same location
  uint32_t bar0;
  __device__ __forceinline__ uint32_t w_full(int s) const { return bar0 + 8u * s; }
  __device__ __forceinline__ uint32_t w_empty(int s) const { return bar0 + 8u * (NSW + s); }
  __device__ __forceinline__ uint32_t kv_full(int s) const { return bar0 + 8u * (2 * NSW + s); }
  __device__ __forceinline__ uint32_t kv_empty(int s) const { return bar0 + 8u * (2 * NSW + NSKV + s); }
  __device__ __forceinline__ uint32_t a_full(int s) const { return bar0 + 8u * (2 * NSW + 2 * NSKV + s); }
  __device__ __forceinline__ uint32_t a_empty(int s) const { return bar0 + 8u * (2 * NSW + 2 * NSKV + NSA + s); }
  __device__ __forceinline__ uint32_t t_full(int s) const { return bar0 + 8u * (2 * NSW + 2 * NSKV + 2 * NSA + s); }
  __device__ __forceinline__ uint32_t t_empty(int s) const { return bar0 + 8u * (2 * NSW + 2 * NSKV + 2 * NSA + 2 + s); }
};

// ---- weight producer (warp 16, lane 0) --------------------------------------------------------------------------------
struct WProd {
  const Ctx& c;
  volatile int* stop;
  uint32_t it = 0;  // ring slots issued so far
  __device__ WProd(const Ctx& c_, volatile int* s) : c(c_), stop(s) {}
  // one CTA tile of a projection: nkb k-blocks of R rows starting at (row n0, k-block kb0)
  __device__ bool tile(const CUtensorMap* map, const GemmCfg& g, int n0, int kb0) {
    for (int kb = 0; kb < g.nkb; kb += g.kbps) {
      const int nk = min(g.kbps, g.nkb - kb), slot = it % NSW;
      if (!mbar_wait_or_stop(c.w_empty(slot), ((it / NSW) & 1) ^ 1, stop)) return false;
      mbar_expect_tx(c.w_full(slot), (uint32_t)nk * g.R * (BK * 2));
      for (int j = 0; j < nk; ++j)
        tma_load_2d(c.base + OFF_W + slot * W_SLOT + j * g.rpad * (BK * 2), map, c.w_full(slot), (kb0 + kb + j) * BK, n0);
      ++it;
    }
    return true;
  }
  __device__ void drain() {  // every load that was issued has landed before the CTA may exit
    const uint32_t first = it > NSW ? it - NSW : 0;
    for (uint32_t i = first; i < it; ++i) fd_wait(c.w_full(i % NSW), (i / NSW) & 1, 11);
  }
};

__device__ void w_producer(const Params& p, const Ctx& c, volatile int* stop) {
  WProd w(c, stop);
  const int cta = blockIdx.x;
  bool ok = true;
  for (int pos = 0; ok && pos + 1 < p.max_length; ++pos) {
    for (int l = 0; ok && l < p.L; ++l) {
      const CUtensorMap* m = p.maps + l * MAPS_PER_LAYER;
      Plan pl = plan_of(p.g_qkv, cta);
      if (pl.active) ok = w.tile(m + M_WQKV, p.g_qkv, pl.n0, pl.kb0);
      pl = plan_of(p.g_dd, cta);
      if (ok && pl.active) ok = w.tile(m + M_WO, p.g_dd, pl.n0, pl.kb0);
      if (ok && pl.active) ok = w.tile(m + M_WQX, p.g_dd, pl.n0, pl.kb0);
      if (ok && pl.active) ok = w.tile(m + M_WOX, p.g_dd, pl.n0, pl.kb0);
      pl = plan_of(p.g_fc1, cta);
      if (ok && pl.active) ok = w.tile(m + M_W1, p.g_fc1, pl.n0, pl.kb0);
      pl = plan_of(p.g_fc2, cta);
      if (ok && pl.active) ok = w.tile(m + M_W2, p.g_fc2, pl.n0, pl.kb0);
    }
    if (ok && pos >= p.n_prompt - 1 && !(p.dbg_skip & 2))
      for (int t = cta; ok && t < p.g_voc.tiles; t += p.G)
        ok = w.tile(p.maps + p.L * MAPS_PER_LAYER + M_VOCAB, p.g_voc, t * p.g_voc.R, 0);
  }
  w.drain();
}

// ---- cross K/V producer (warp 17, lane 0) ----------------------------------------------------------------------------
__device__ void kv_producer(const Params& p, const Ctx& c, volatile int* stop) {
  uint32_t it = 0;
  const int cta = blockIdx.x, pairs = p.B * p.H;
  const uint32_t bytes = (uint32_t)p.TR * (HD * 2);
  bool ok = !(p.dbg_skip & 1);
  for (int pos = 0; ok && pos + 1 < p.max_length; ++pos)
    for (int l = 0; ok && l < p.L; ++l) {
      const CUtensorMap* m = p.maps + l * MAPS_PER_LAYER + M_XKV;
      for (int pr = cta; ok && pr < pairs; pr += p.G) {
        const int b = pr / p.H, h = pr % p.H;
        for (int t = 0; ok && t < p.n_kv_tiles; ++t)
          for (int kv = 0; kv < 2; ++kv) {  // tile t of K, then tile t of V: the consumer folds them in together
            const int slot = it % NSKV;
            if (!mbar_wait_or_stop(c.kv_empty(slot), ((it / NSKV) & 1) ^ 1, stop)) { ok = false; break; }
            mbar_expect_tx(c.kv_full(slot), bytes);
            tma_load_2d(c.base + OFF_KV + slot * KV_SLOT, m, c.kv_full(slot), kv * p.d + h * HD, b * p.S + t * p.TR);
            ++it;
          }
      }
    }
  const uint32_t first = it > NSKV ? it - NSKV : 0;
  for (uint32_t i = first; i < it; ++i) fd_wait(c.kv_full(i % NSKV), (i / NSKV) & 1, 12);
}

// ---- per-thread pipeline counters of the main (barrier-synchronised) warps -----------------------------------------------
struct Counters {
  uint32_t w_ct = 0;   // weight ring slots consumed (MMA thread)
  uint32_t a_it = 0;   // activation k-blocks issued (A thread)
  uint32_t a_ct = 0;   // activation k-blocks consumed (MMA thread)
  uint32_t t_ct = 0;   // accumulator tiles (MMA thread and epilogue warps count alike)
  uint32_t kv_ct = 0;  // cross K/V tiles consumed (compute warps)
};

__device__ __forceinline__ void a_thread_tile(const Ctx& c, Counters& k, const CUtensorMap* amap, int kb0, int nkb) {
  for (int kb = 0; kb < nkb; ++kb) {
    const int slot = k.a_it % NSA;
    fd_wait(c.a_empty(slot), ((k.a_it / NSA) & 1) ^ 1, 3);
    mbar_expect_tx(c.a_full(slot), A_SLOT);
    tma_load_2d(c.base + OFF_A + slot * A_SLOT, amap, c.a_full(slot), (kb0 + kb) * BK, 0);
    ++k.a_it;
  }
}

__device__ __forceinline__ void mma_thread_tile(const Ctx& c, Counters& k, uint32_t tmem_base, const GemmCfg& g) {
  const uint32_t buf = k.t_ct & 1;
  fd_wait(c.t_empty(buf), ((k.t_ct >> 1) & 1) ^ 1, 4);
  tc_fence_after();
  const uint32_t d_tmem = tmem_base + buf * NB;
  for (int kb = 0; kb < g.nkb; ++kb) {
    const int j = kb % g.kbps, wslot = k.w_ct % NSW, aslot = k.a_ct % NSA;
    if (j == 0) fd_wait(c.w_full(wslot), (k.w_ct / NSW) & 1, 5);
    fd_wait(c.a_full(aslot), (k.a_ct / NSA) & 1, 6);
    tc_fence_after();
    const uint64_t dw = make_desc(c.base + OFF_W + wslot * W_SLOT + j * g.rpad * (BK * 2));
    const uint64_t da = make_desc(c.base + OFF_A + aslot * A_SLOT);
#pragma unroll
    for (int q = 0; q < BK / UMMA_K; ++q) umma_f16(d_tmem, dw + 2 * q, da + 2 * q, IDESC, (kb | q) != 0);
    umma_commit(c.a_empty(aslot));
    ++k.a_ct;
    if (j == g.kbps - 1 || kb == g.nkb - 1) {
      umma_commit(c.w_empty(wslot));
      ++k.w_ct;
    }
  }
  umma_commit(c.t_full(buf));
  ++k.t_ct;
}

// block-wide reductions over the 512 compute threads (barrier id 2); s_red: >= 2 * N_CWARPS floats of scratch
__device__ __forceinline__ float block_sum(float v, float* s_red, int warp, int lane) {
  v = warp_sum(v);
  named_sync(2, N_CTHREADS);  // previous use of s_red finished
  if (lane == 0) s_red[warp] = v;
  named_sync(2, N_CTHREADS);
  float t = 0.0f;
#pragma unroll
  for (int i = 0; i < N_CWARPS; ++i) t += s_red[i];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* s_red, int warp, int lane) {
  v = warp_max(v);
  named_sync(2, N_CTHREADS);
  if (lane == 0) s_red[warp] = v;
  named_sync(2, N_CTHREADS);
  float t = s_red[0];
#pragma unroll
  for (int i = 1; i < N_CWARPS; ++i) t = fmaxf(t, s_red[i]);
  return t;
}

enum RowKind { ROW_EMBED, ROW_RES };

// Row-owner phase for batch row b (one CTA): x[b] = embedding, or x[b] += bias + sum of the K-group partials of the
// preceding projection; then LayerNorm(x[b]) -> da[b] (bf16), the next projection's operand.
__device__ void row_phase(const Params& p, float* scr, int b, int kind, int pos, const float* bias, int n_part,
                          const float* ln_w, const float* ln_b, int tid, int warp, int lane) {
  const int d = p.d, c = tid * 4;
  const bool on = c < d;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (on) {
    if (kind == ROW_EMBED) {
      int tok = p.tokens[(size_t)b * p.ld_tokens + pos];
      tok = min(max(tok, 0), p.V - 1);
      const float4 e = ld4(p.tok_embed + (size_t)tok * d + c);
      const float4 q = *reinterpret_cast<const float4*>(p.dec_pos + (size_t)pos * d + c);
      v = make_float4(e.x + q.x, e.y + q.y, e.z + q.z, e.w + q.w);
    } else {
      v = *reinterpret_cast<const float4*>(p.x + (size_t)b * d + c);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c));
      float4 acc = __ldcg(reinterpret_cast<const float4*>(p.part + (size_t)b * d + c));
      for (int s = 1; s < n_part; ++s) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(p.part + ((size_t)s * NB + b) * d + c));
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
      v.x += acc.x + bb.x; v.y += acc.y + bb.y; v.z += acc.z + bb.z; v.w += acc.w + bb.w;
    }
    *reinterpret_cast<float4*>(p.x + (size_t)b * d + c) = v;
  }
  if (!ln_w) return;
  const float mean = block_sum(on ? (v.x + v.y) + (v.z + v.w) : 0.0f, scr, warp, lane) / (float)d;
  float sq = 0.0f;
  if (on) {
    const float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
    sq = (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
  }
  const float rstd = rsqrtf(block_sum(sq, scr + N_CWARPS, warp, lane) / (float)d + 1e-5f);
  if (on) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(ln_w + c)), be = __ldg(reinterpret_cast<const float4*>(ln_b + c));
    st4(p.da + (size_t)b * d + c, make_float4((v.x - mean) * rstd * g.x + be.x, (v.y - mean) * rstd * g.y + be.y,
                                             (v.z - mean) * rstd * g.z + be.z, (v.w - mean) * rstd * g.w + be.w));
  }
}

enum EpiKind { EPI_PART, EPI_PARTQ, EPI_QKV, EPI_FC1 };

// Compute-warp half of one projection tile: accumulator (lane = weight row, column = batch row) -> transposed through
// shared memory (the activation ring: every k-block of the tile has been consumed when the accumulator is complete, and
// the next activation loads are issued after the phase barrier) -> coalesced global stores.
__device__ void epilogue_tile(const Params& p, const Ctx& c, Counters& k, uint32_t tmem_base, float* stage,
                              const GemmCfg& g, const Plan& pl, int kind, const float* bias, int tid, int warp, int lane) {
  const uint32_t buf = k.t_ct & 1;
  const int q = warp & 3, cg = warp >> 2, RP = g.rpad + 1;
  fd_wait(c.t_full(buf), (k.t_ct >> 1) & 1, 7);
  tc_fence_after();
  if (32 * q < g.R) {
    uint32_t r[16];
    tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + buf * NB + 16 * cg, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int row = 32 * q + lane;
    if (row < g.R) {
#pragma unroll
      for (int j = 0; j < 16; ++j) stage[(16 * cg + j) * RP + row] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(c.t_empty(buf));
  ++k.t_ct;
  named_sync(2, N_CTHREADS);
  const int items = p.B * g.R;
  for (int i = tid; i < items; i += N_CTHREADS) {
    const int b = i / g.R, f = i % g.R, n = pl.n0 + f;
    if (n >= g.N) continue;
    float v = stage[b * RP + f];
    if (kind == EPI_PART) {
      p.part[((size_t)pl.kg * NB + b) * g.N + n] = v;
    } else if (kind == EPI_PARTQ) {
      p.partq[((size_t)pl.kg * NB + b) * g.N + n] = v;
    } else if (kind == EPI_QKV) {
      p.partqkv[((size_t)pl.kg * NB + b) * g.N + n] = v;
    } else {
      p.dh[(size_t)b * g.N + n] = __float2bfloat16_rn(gelu_as(v + __ldg(bias + n)));
    }
  }
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float LOG2E = 1.4426950408889634f;

// Running soft-max state of one 8-lane sub-group (each lane: 8 of the 64 output dims): fold in one key / value row.
// Scores are in log2 units (q is pre-multiplied by log2 e), so exp is a bare ex2.
__device__ __forceinline__ void online_row(float& m, float& l, float* o, float s, const float* vf) {
  const float mn = fmaxf(m, s), corr = ex2f(m - mn), pe = ex2f(s - mn);
  l = fmaf(l, corr, pe);
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = fmaf(o[e], corr, pe * vf[e]);
  m = mn;
}
// merge the states of the 4 sub-groups of a warp (lanes l8, l8 + 8, l8 + 16, l8 + 24 hold the same 8 dims)
__device__ __forceinline__ void merge_subgroups(float& m, float& l, float* o) {
#pragma unroll
  for (int off = 8; off <= 16; off <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, off), l2 = __shfl_xor_sync(0xffffffffu, l, off);
    const float mn = fmaxf(m, m2);
    const float c1 = mn == -INFINITY ? 0.0f : ex2f(m - mn), c2 = mn == -INFINITY ? 0.0f : ex2f(m2 - mn);
    l = l * c1 + l2 * c2;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = o[e] * c1 + __shfl_xor_sync(0xffffffffu, o[e], off) * c2;
    m = mn;
  }
}

// Self-attention for one decoder position, ONE WARP per (batch, head) pair, no block-level synchronisation: the warp adds
// up the K-group partials of the fused q | k | v projection (+ bias), appends the new k / v row to the preallocated
// cache, and runs an online soft-max over the <= 448 cached rows, 4 rows per step (8 lanes per row, 16-byte loads),
// 4 steps in flight.  Rows [row0, row0 + nrows) of the batch; `nw` warps of this CTA take part (wi = 0 .. nw - 1).
__device__ void self_attn_phase(const Params& p, const LayerP& L, int pos, int row0, int nrows, int wi, int nw, int lane) {
  const int sub = lane >> 3, l8 = lane & 7, d = p.d, n = pos + 1, n_part = p.g_qkv.S;
  const int pairs = nrows * p.H;
  for (int pr = blockIdx.x * nw + wi; pr < pairs; pr += p.G * nw) {
    const int b = row0 + pr / p.H, h = pr % p.H;
    // this lane's 8 features of q (all sub-groups), of k (sub-group 0) or of v (sub-group 1)
    float qf[8], nf[8];
    const int qcol = h * HD + l8 * 8, ncol = (sub == 0 ? d : 2 * d) + qcol;
#pragma unroll
    for (int e = 0; e < 8; ++e) { qf[e] = __ldg(L.bqkv + qcol + e); nf[e] = sub < 2 ? __ldg(L.bqkv + ncol + e) : 0.0f; }
    for (int s = 0; s < n_part; ++s) {
      const float* row = p.partqkv + ((size_t)s * NB + b) * 3 * d;
      const float4 a0 = __ldcg(reinterpret_cast<const float4*>(row + qcol)), a1 = __ldcg(reinterpret_cast<const float4*>(row + qcol + 4));
      qf[0] += a0.x; qf[1] += a0.y; qf[2] += a0.z; qf[3] += a0.w; qf[4] += a1.x; qf[5] += a1.y; qf[6] += a1.z; qf[7] += a1.w;
      if (sub < 2) {
        const float4 c0 = __ldcg(reinterpret_cast<const float4*>(row + ncol)), c1 = __ldcg(reinterpret_cast<const float4*>(row + ncol + 4));
        nf[0] += c0.x; nf[1] += c0.y; nf[2] += c0.z; nf[3] += c0.w; nf[4] += c1.x; nf[5] += c1.y; nf[6] += c1.z; nf[7] += c1.w;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) qf[e] *= LOG2E;
    bf16* kp = L.self_k + ((size_t)b * p.H + h) * p.MT * HD;
    bf16* vp = L.self_v + ((size_t)b * p.H + h) * p.MT * HD;
    if (sub < 2) {
      uint4 pk;
      pk.x = pack_bf16(nf[0], nf[1]); pk.y = pack_bf16(nf[2], nf[3]); pk.z = pack_bf16(nf[4], nf[5]); pk.w = pack_bf16(nf[6], nf[7]);
      *reinterpret_cast<uint4*>((sub == 0 ? kp : vp) + (size_t)pos * HD + l8 * 8) = pk;
    }
    __syncwarp();  // the appended row is visible to the whole warp
    float m = -INFINITY, l = 0.0f, o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.0f;
    constexpr int SU = 3;  // row steps in flight per sub-group (6 x 16-byte loads per lane)
    for (int j0 = 0; j0 < n; j0 += 4 * SU) {
      uint4 kr[SU], vr[SU];
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        const int j = j0 + u * 4 + sub;
        if (j < n) {
          kr[u] = *reinterpret_cast<const uint4*>(kp + (size_t)j * HD + l8 * 8);
          vr[u] = *reinterpret_cast<const uint4*>(vp + (size_t)j * HD + l8 * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        const int j = j0 + u * 4 + sub;
        float kf[8], vf[8], acc = 0.0f;
        if (j < n) {
          unpack8(kr[u], kf);
          unpack8(vr[u], vf);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc = fmaf(qf[e], kf[e], acc);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (j < n) online_row(m, l, o, acc, vf);
      }
    }
    merge_subgroups(m, l, o);
    if (sub == 0) {
      const float inv = 1.0f / l;
      uint4 pk;
      pk.x = pack_bf16(o[0] * inv, o[1] * inv); pk.y = pack_bf16(o[2] * inv, o[3] * inv);
      pk.z = pack_bf16(o[4] * inv, o[5] * inv); pk.w = pack_bf16(o[6] * inv, o[7] * inv);
      *reinterpret_cast<uint4*>(p.dattn + (size_t)b * d + h * HD + l8 * 8) = pk;
    }
  }
}

// Cross-attention for one decoder position over rows [row0, row0 + nrows) of the batch: this CTA's (batch, head) pairs
// one after the other, `nw` warps (wi = 0 .. nw - 1, named barrier `bar`) on the K / V tiles the producer warp streams
// through the ring (tile t of K, then tile t of V).  Online soft-max per 8-lane sub-group — no block-wide pass over the
// scores — and one merge of the warps' states per pair.
__device__ void cross_attn_phase(const Params& p, const Ctx& c, Counters& k, const LayerP& L, float* scr, bf16* out,
                                 int row0, int nrows, int wi, int nw, int lane, int bar) {
  float* s_q = scr;              // [64]
  float* s_m = s_q + HD;         // [nw]
  float* s_l = s_m + N_CWARPS;   // [nw]
  float* s_o = s_l + N_CWARPS;   // [nw][64]
  const int xtid = wi * 32 + lane, nthr = nw * 32;
  const int pairs = nrows * p.H, d = p.d, S = p.S, sub = lane >> 3, l8 = lane & 7;
  const int n_part = p.g_dd.S, swz = p.kv_sw128 ? 7 : 0;
  if (p.dbg_skip & 1) {
    for (int i = xtid; i < p.B * d; i += nthr) if (blockIdx.x == 0) out[i] = __float2bfloat16_rn(0.0f);
    return;
  }
  for (int pr = blockIdx.x; pr < pairs; pr += p.G) {
    const int b = row0 + pr / p.H, h = pr % p.H;
    if (xtid < HD) {  // q = bias + sum of the cross-q projection's K-group partials (fixed order), in log2 units
      float q = __ldg(L.bq_x + h * HD + xtid);
      for (int s = 0; s < n_part; ++s) q += __ldcg(p.partq + ((size_t)s * NB + b) * d + h * HD + xtid);
      s_q[xtid] = q * LOG2E;
    }
    named_sync(bar, nthr);
    float qf[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) qf[e] = s_q[l8 * 8 + e];
    float m = -INFINITY, l = 0.0f, o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.0f;
    for (int t = 0; t < p.n_kv_tiles; ++t) {
      const int slot_k = k.kv_ct % NSKV, slot_v = (k.kv_ct + 1) % NSKV;
      fd_wait(c.kv_full(slot_k), (k.kv_ct / NSKV) & 1, 8);
      fd_wait(c.kv_full(slot_v), ((k.kv_ct + 1) / NSKV) & 1, 9);
      const uint8_t* tk = c.gen + OFF_KV + slot_k * KV_SLOT;
      const uint8_t* tv = c.gen + OFF_KV + slot_v * KV_SLOT;
      const int rows = min(p.TR, S - t * p.TR);
      for (int r = wi * 4 + sub; r < rows; r += nw * 4) {
        const int off = r * (HD * 2) + ((l8 ^ (r & swz)) * 16);
        float kf[8], vf[8];
        unpack8(*reinterpret_cast<const uint4*>(tk + off), kf);
        unpack8(*reinterpret_cast<const uint4*>(tv + off), vf);
        float acc = 0.0f;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc = fmaf(qf[e], kf[e], acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        online_row(m, l, o, acc, vf);
      }
      __syncwarp();
      if (lane == 0) { mbar_arrive(c.kv_empty(slot_k)); mbar_arrive(c.kv_empty(slot_v)); }
      k.kv_ct += 2;
    }
    merge_subgroups(m, l, o);
    if (sub == 0) {
      if (l8 == 0) { s_m[wi] = m; s_l[wi] = l; }
#pragma unroll
      for (int e = 0; e < 8; ++e) s_o[wi * HD + l8 * 8 + e] = o[e];
    }
    named_sync(bar, nthr);
    if (xtid < HD) {
      float mx = s_m[0];
      for (int w = 1; w < nw; ++w) mx = fmaxf(mx, s_m[w]);
      float lt = 0.0f, acc = 0.0f;
      for (int w = 0; w < nw; ++w) {
        const float cw = s_m[w] == -INFINITY ? 0.0f : ex2f(s_m[w] - mx);
        lt = fmaf(s_l[w], cw, lt);
        acc = fmaf(s_o[w * HD + xtid], cw, acc);
      }
      out[(size_t)b * d + h * HD + xtid] = __float2bfloat16_rn(acc / lt);
    }
    named_sync(bar, nthr);  // scratch free for the next pair
  }
}

// Row state of the timestamp rules, re-derived from the token history (exactly what the HF processor derives from
// input_ids[k, begin_index:]): bit0 at_begin, bit1 last token is a timestamp, bit2 penultimate is a timestamp (or fewer
// than two sampled), bit3 any timestamp so far; bound = first timestamp id still allowed.
__device__ __forceinline__ void row_state(const Params& p, int b, int pos, int* st, int* bound) {
  const int* trow = p.tokens + (size_t)b * p.ld_tokens;
  const int n = pos + 1 - p.n_prompt, tb = p.rules.ts_begin;
  const int last_ts = n >= 1 && trow[pos] >= tb;
  const int pen_ts = n < 2 || trow[pos - 1] >= tb;
  int has_ts = 0, ts_last = 0;
  for (int j = pos; j >= p.n_prompt; --j)
    if (trow[j] >= tb) {
      has_ts = 1;
      ts_last = trow[j];
      break;
    }
  *st = (n == 0 ? 1 : 0) | (last_ts << 1) | (pen_ts << 2) | (has_ts << 3);
  *bound = (last_ts && !pen_ts) ? ts_last : ts_last + 1;
}
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

// Compute-warp half of one vocabulary tile: processors + arg-max partials straight from the accumulator.  Warp (q, cg)
// owns vocabulary rows n0 + 32 q .. + 31 and batch columns 16 cg .. + 15; it emits one partial per batch column.
__device__ void vocab_epilogue(const Params& p, const Ctx& c, Counters& k, uint32_t tmem_base, const int* s_st,
                               const int* s_bound, int tile, int warp, int lane) {
  const uint32_t buf = k.t_ct & 1;
  const int q = warp & 3, cg = warp >> 2, tb = p.rules.ts_begin;
  fd_wait(c.t_full(buf), (k.t_ct >> 1) & 1, 10);
  tc_fence_after();
  uint32_t r[16];
  tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + buf * NB + 16 * cg, r);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(c.t_empty(buf));
  ++k.t_ct;
  const int n0 = tile * p.g_voc.R, rowi = 32 * q + lane, v = n0 + rowi;
  const bool valid = rowi < p.g_voc.R && v < p.V;
  const unsigned f = valid ? p.flags[v] : 1u;
  const int lo = n0 + 32 * q, hi = lo + 31;
  const bool has_text = lo < tb, has_ts = hi >= tb;
  Best my_t = {-INFINITY, p.V}, my_s = {-INFINITY, p.V};
  float my_sum = 0.0f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int b = 16 * cg + j;
    if (b >= p.B) break;
    const float xv = __uint_as_float(r[j]);
    const bool ok = valid && !token_masked(p.rules, p.return_ts, v, f, s_st[b], s_bound[b]);
    Best bt = {-INFINITY, p.V}, bs = {-INFINITY, p.V};
    float sum = 0.0f;
    if (has_text) {
      Best cnd = {(ok && v < tb) ? xv : -INFINITY, (ok && v < tb) ? v : p.V};
      bt = warp_best(cnd);
    }
    if (has_ts) {
      const bool on = ok && v >= tb;
      Best cnd = {on ? xv : -INFINITY, on ? v : p.V};
      bs = warp_best(cnd);
      if (p.return_ts && bs.v > -INFINITY) sum = warp_sum(on ? __expf(xv - bs.v) : 0.0f);
    }
    if (lane == j) { my_t = bt; my_s = bs; my_sum = sum; }
  }
  const int b = 16 * cg + lane;
  if (lane < 16 && b < p.B) {
    float* o = p.vpart + ((size_t)b * (p.g_voc.tiles * 4) + tile * 4 + q) * VP_WORDS;
    o[0] = my_t.v; o[1] = __int_as_float(my_t.i); o[2] = my_s.v; o[3] = __int_as_float(my_s.i); o[4] = my_sum;
  }
}

// Row-owner token pick for batch row b: combine the vocabulary partials, apply the "timestamps outweigh text" rule, write
// the next token and the finished flag (GenerationMixin._sample tail).
__device__ void combine_phase(const Params& p, float* scr, int b, int pos, int tid, int warp, int lane) {
  const int np = p.g_voc.tiles * 4;
  const float* vp = p.vpart + (size_t)b * np * VP_WORDS;
  Best bt = {-INFINITY, p.V}, bs = {-INFINITY, p.V};
  for (int i = tid; i < np; i += N_CTHREADS) {
    const float* o = vp + (size_t)i * VP_WORDS;
    Best t = {__ldcg(o), __float_as_int(__ldcg(o + 1))}, s = {__ldcg(o + 2), __float_as_int(__ldcg(o + 3))};
    bt = better(bt, t);
    bs = better(bs, s);
  }
  bt = warp_best(bt);
  bs = warp_best(bs);
  Best* s_b = reinterpret_cast<Best*>(scr + 4 * N_CWARPS);  // [2][N_CWARPS]
  named_sync(2, N_CTHREADS);
  if (lane == 0) { s_b[warp] = bt; s_b[N_CWARPS + warp] = bs; }
  named_sync(2, N_CTHREADS);
  bt = s_b[0];
  bs = s_b[N_CWARPS];
  for (int w = 1; w < N_CWARPS; ++w) { bt = better(bt, s_b[w]); bs = better(bs, s_b[N_CWARPS + w]); }
  int choice;
  if (p.return_ts) {
    float part = 0.0f;
    if (bs.v > -INFINITY)
      for (int i = tid; i < np; i += N_CTHREADS) {
        const float* o = vp + (size_t)i * VP_WORDS;
        const float m = __ldcg(o + 2);
        if (m > -INFINITY) part += __ldcg(o + 4) * __expf(m - bs.v);
      }
    const float tot = block_sum(part, scr, warp, lane);
    const float lse = (bs.v > -INFINITY) ? bs.v + logf(tot) : -INFINITY;
    choice = (lse > bt.v) ? bs.i : ((bt.v >= bs.v) ? bt.i : bs.i);
  } else {
    choice = (bt.v >= bs.v) ? bt.i : bs.i;
  }
  if (tid == 0) {
    if (choice >= p.V) choice = 0;
    const int fin = p.finished[b];
    const int next = fin ? p.rules.pad : choice;
    p.tokens[(size_t)b * p.ld_tokens + pos + 1] = next;
    if (next == p.rules.eos) p.finished[b] = 1;
  }
  named_sync(2, N_CTHREADS);  // the new token is visible to the CTA (the embedding of the next position reads it)
}

__global__ void __launch_bounds__(THREADS, 1) dec_fused_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  Ctx c;
  c.base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  c.gen = smem_raw + (c.base - smem_u32(smem_raw));
  c.bar0 = c.base + OFF_BAR;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(c.gen + OFF_MISC);
  volatile int* s_stop = reinterpret_cast<volatile int*>(c.gen + OFF_MISC + 4);
  volatile int* s_allfin = reinterpret_cast<volatile int*>(c.gen + OFF_MISC + 8);
  float* scr = reinterpret_cast<float*>(c.gen + OFF_SCR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, cta = blockIdx.x;

  if (tid == 0) {
    g_dbg = p.dbg;
    for (int s = 0; s < NSW; ++s) { mbar_init(c.w_full(s), 1); mbar_init(c.w_empty(s), 1); }
    for (int s = 0; s < NSKV; ++s) { mbar_init(c.kv_full(s), 1); mbar_init(c.kv_empty(s), N_CWARPS); }
    for (int s = 0; s < NSA; ++s) { mbar_init(c.a_full(s), 1); mbar_init(c.a_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(c.t_full(s), 1); mbar_init(c.t_empty(s), N_CWARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *s_stop = 0;
    *s_allfin = 0;
  }
  if (warp == MMA_WARP) tmem_alloc(c.base + OFF_MISC, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == W_WARP) {
    if (lane == 0) w_producer(p, c, s_stop);
  } else if (warp == KV_WARP) {
    if (lane == 0) kv_producer(p, c, s_stop);
  } else {
    // ---- main warps: phases separated by grid barriers -----------------------------------------------------------------
    Counters k;
    unsigned bar_target = 0;
    const bool is_a = warp == A_WARP && lane == 0, is_mma = warp == MMA_WARP && lane == 0, is_c = warp < N_CWARPS;
    const CUtensorMap* gmaps = p.maps + p.L * MAPS_PER_LAYER;

    int crumb = 0;
    // bring-up timing (KW_FUSED_DEBUG): CTA 0 / thread 0 splits its time per phase kind into work (barrier exit -> next
    // barrier entry) and barrier (entry -> exit); written behind the breadcrumbs at kernel end
    constexpr int NPK = 11;
    unsigned long long acc_work[NPK], acc_bar[NPK], t_mark = 0;
    unsigned acc_n[NPK];
#ifndef KW_FD_TIMING
#define KW_FD_TIMING 0  // build with KW_NVCC_EXTRA=-DKW_FD_TIMING=1 for the per-phase timing table (costs registers)
#endif
    const bool timing = KW_FD_TIMING && p.dbg != nullptr && tid == 0 && cta == 0;
    if (timing) {
      for (int i = 0; i < NPK; ++i) { acc_work[i] = 0; acc_bar[i] = 0; acc_n[i] = 0; }
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_mark));
    }
    auto phase_sync = [&](int pk) {
      ++crumb;
      if (tid == 0 && p.dbg) p.dbg[cta * 16] = (unsigned)crumb;
      unsigned long long t_in = 0;
      if (timing) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_in));
        acc_work[pk] += t_in - t_mark;
        acc_n[pk] += 1;
      }
      fence_async_proxy();  // this thread's global writes precede later TMA (async-proxy) reads by any CTA
      named_sync(1, SYNC_THREADS);
      if (tid == 0) {
        bar_target += gridDim.x;
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.bar_counter) : "memory");
        unsigned v;
        const long long t0 = clock64();
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.bar_counter) : "memory");
          if (clock64() - t0 > 8000000000LL) {
            if (p.dbg) {
              p.dbg[cta * 16 + 9] = 0xC0000000u | v;
              __threadfence_system();
            }
            printf("kwb200 decode_fused: grid barrier timed out (block %d, %u of %u)\n", blockIdx.x, v, bar_target);
            __trap();
          }
        } while (v < bar_target);
        __threadfence();
      }
      named_sync(1, SYNC_THREADS);
      fence_async_proxy();
      if (timing) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_mark));
        acc_bar[pk] += t_mark - t_in;
      }
    };
    // one projection phase; amap = activation operand, kind / bias = epilogue
    auto gemm_phase = [&](const GemmCfg& g, const CUtensorMap* amap, int kind, const float* bias, int pk) {
      const Plan pl = plan_of(g, cta);
      if (pl.active) {
        if (is_a) a_thread_tile(c, k, amap, pl.kb0, g.nkb);
        else if (is_mma) mma_thread_tile(c, k, tmem_base, g);
        else if (is_c) epilogue_tile(p, c, k, tmem_base, reinterpret_cast<float*>(c.gen + OFF_A), g, pl, kind, bias, tid, warp, lane);
      }
      __syncwarp();
      phase_sync(pk);
    };

    // prologue: embedding + first LayerNorm of position 0
    if (is_c && cta < p.B) row_phase(p, scr, cta, ROW_EMBED, 0, nullptr, 0, p.layers[0].ln1_w, p.layers[0].ln1_b, tid, warp, lane);
    phase_sync(0);

    int steps = 0;
    for (int pos = 0; pos + 1 < p.max_length; ++pos) {
      const bool sample = pos >= p.n_prompt - 1;
      for (int l = 0; l < p.L; ++l) {
        const LayerP& L = p.layers[l];
        const CUtensorMap* lm = p.maps + l * MAPS_PER_LAYER;
        (void)lm;
        if (l > 0) {
          if (is_c && cta < p.B)
            row_phase(p, scr, cta, ROW_RES, pos, p.layers[l - 1].b2, p.g_fc2.S, L.ln1_w, L.ln1_b, tid, warp, lane);
          phase_sync(0);
        }
        gemm_phase(p.g_qkv, gmaps + M_DA, EPI_QKV, L.bqkv, 1);
        if (is_c) self_attn_phase(p, L, pos, 0, p.B, warp, N_CWARPS, lane);
        phase_sync(2);
        gemm_phase(p.g_dd, gmaps + M_DATTN, EPI_PART, nullptr, 3);
        if (is_c && cta < p.B) row_phase(p, scr, cta, ROW_RES, pos, L.bo, p.g_dd.S, L.lnx_w, L.lnx_b, tid, warp, lane);
        phase_sync(0);
        gemm_phase(p.g_dd, gmaps + M_DA, EPI_PARTQ, nullptr, 4);
        if (is_c) cross_attn_phase(p, c, k, L, scr, p.dattn, 0, p.B, warp, N_CWARPS, lane, 2);
        phase_sync(5);
        gemm_phase(p.g_dd, gmaps + M_DATTN, EPI_PART, nullptr, 6);
        if (is_c && cta < p.B) row_phase(p, scr, cta, ROW_RES, pos, L.bo_x, p.g_dd.S, L.ln3_w, L.ln3_b, tid, warp, lane);
        phase_sync(0);
        gemm_phase(p.g_fc1, gmaps + M_DA, EPI_FC1, L.b1, 7);
        gemm_phase(p.g_fc2, gmaps + M_DH, EPI_PART, nullptr, 8);
      }
      ++steps;
      if (sample) {
        if (is_c && cta < p.B)
          row_phase(p, scr, cta, ROW_RES, pos, p.layers[p.L - 1].b2, p.g_fc2.S, p.lnf_w, p.lnf_b, tid, warp, lane);
        phase_sync(0);
        // vocabulary projection: tiles cta, cta + G, ...; per-row rule state first (every CTA, from the token history)
        int* s_st = reinterpret_cast<int*>(scr);
        int* s_bound = s_st + NB;
        if (is_c && tid < p.B) {
          int st, bd;
          row_state(p, tid, pos, &st, &bd);
          s_st[tid] = st;
          s_bound[tid] = bd;
        }
        if (is_c) named_sync(2, N_CTHREADS);
        for (int t = cta; t < p.g_voc.tiles && !(p.dbg_skip & 2); t += p.G) {
          if (is_a) a_thread_tile(c, k, gmaps + M_DA, 0, p.g_voc.nkb);
          else if (is_mma) mma_thread_tile(c, k, tmem_base, p.g_voc);
          else if (is_c) vocab_epilogue(p, c, k, tmem_base, s_st, s_bound, t, warp, lane);
        }
        __syncwarp();
        phase_sync(9);
        if (is_c && cta < p.B) {
          combine_phase(p, scr, cta, pos, tid, warp, lane);
          if (pos + 2 < p.max_length)
            row_phase(p, scr, cta, ROW_EMBED, pos + 1, nullptr, 0, p.layers[0].ln1_w, p.layers[0].ln1_b, tid, warp, lane);
        }
        phase_sync(10);
        // all rows finished -> the pass is over (GenerationMixin stops when unfinished_sequences.max() == 0)
        if (tid == 0) {
          int all = 1;
          for (int b = 0; b < p.B; ++b) all &= (__ldcg(p.finished + b) != 0);
          *s_allfin = all;
        }
        named_sync(1, SYNC_THREADS);
        if (*s_allfin) break;
      } else {
        if (is_c && cta < p.B)
          row_phase(p, scr, cta, ROW_EMBED, pos + 1, nullptr, 0, p.layers[0].ln1_w, p.layers[0].ln1_b, tid, warp, lane);
        phase_sync(10);
      }
    }
    if (timing)
      for (int i = 0; i < NPK; ++i) {
        p.dbg[gridDim.x * 16 + i * 3] = (unsigned)(acc_work[i] / 1000);
        p.dbg[gridDim.x * 16 + i * 3 + 1] = (unsigned)(acc_bar[i] / 1000);
        p.dbg[gridDim.x * 16 + i * 3 + 2] = acc_n[i];
      }
    if (tid == 0) {
      *s_stop = 1;
      if (cta == 0) *p.steps_out = steps;
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();  // producers have drained their rings, every MMA was consumed by an epilogue
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace fd

// ---- host side -------------------------------------------------------------------------------------------------------
struct FusedDecode {
  fd::Params p;
  CUtensorMap* d_maps = nullptr;
  fd::LayerP* d_layers = nullptr;
  float* part = nullptr;
  float* partq = nullptr;
  float* partqkv = nullptr;
  float* vpart = nullptr;
  unsigned* bar_counter = nullptr;
  int* steps_dev = nullptr;
  unsigned* dbg_host = nullptr;  // mapped pinned memory, 16 words per CTA
  int n_sm = 0;
  bool ok = false;
};

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

static bool make_cfg(fd::GemmCfg& g, int N, int K, int G, bool partial) {
  if (K % fd::BK) return false;
  const int kb = K / fd::BK;
  g.N = N;
  g.K = K;
  // K-groups: every CTA should see at most ~5-10 k-blocks, so that its activation operand arrives in one TMA round trip
  g.S = 1;
  if (partial) {
    if (kb >= 80 && kb % 8 == 0 && G >= 16) g.S = 8;
    else if (kb >= 8 && kb % 4 == 0 && G >= 8) g.S = 4;
  }
  g.n_groups = G / g.S;
  g.R = (N + g.n_groups - 1) / g.n_groups;
  g.rpad = round_up(g.R, 8);
  g.kbps = std::min(128 / g.rpad, 8);
  g.nkb = kb / g.S;
  g.tiles = 0;
  // the epilogue transposes [batch 64][rpad + 1] floats through the activation ring (idle once the tile's MMAs are done)
  return g.R <= 128 && g.kbps >= 1 && (size_t)(g.rpad + 1) * fd::NB * sizeof(float) <= (size_t)fd::NSA * fd::A_SLOT;
}

void fused_decode_destroy(kw_model* m) {
  FusedDecode* f = m->fused;
  if (!f) return;
  cudaFree(f->d_maps);
  cudaFree(f->d_layers);
  cudaFree(f->part);
  cudaFree(f->partq);
  cudaFree(f->partqkv);
  cudaFree(f->vpart);
  cudaFree(f->bar_counter);
  cudaFree(f->steps_dev);
  if (f->dbg_host) cudaFreeHost(f->dbg_host);
  delete f;
  m->fused = nullptr;
}

static int map2d(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows, bool swz = true) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)fd::BK, (cuuint32_t)box_rows};
  return tc::make_map_bf16(map, ptr, 2, gdim, gstride, box, swz);
}

// Build (once per handle) the tensor maps, layer table and scratch of the fused decode kernel.  Returns KW_OK, or
// KW_ERR_UNSUPPORTED when the model shape does not fit (the caller then keeps the kernel-per-op schedule).
int fused_decode_prepare(kw_model* m) {
  if (m->fused) return m->fused->ok ? KW_OK : KW_ERR_UNSUPPORTED;
  FusedDecode* f = new FusedDecode();
  m->fused = f;
  const kw_config& c = m->cfg;
  if (m->t != KW_BF16 || c.d_model % 64 || c.ffn_dim % 64 || c.d_model > 2048 || c.max_batch > fd::NB ||
      c.max_target_pos > fd::MAX_T || c.max_source_pos > fd::MAX_S) {
    set_error("decode_fused: model shape / dtype outside the fused kernel's envelope");
    return KW_ERR_UNSUPPORTED;
  }
  int dev = 0, coop = 0;
  KW_CUDA_OK(cudaGetDevice(&dev));
  KW_CUDA_OK(cudaDeviceGetAttribute(&f->n_sm, cudaDevAttrMultiProcessorCount, dev));
  KW_CUDA_OK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) {
    set_error("decode_fused: device has no cooperative launch");
    return KW_ERR_UNSUPPORTED;
  }
  KW_CUDA_OK(cudaFuncSetAttribute(fd::dec_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fd::SMEM_BYTES));
  int per_sm = 0;
  KW_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fd::dec_fused_kernel, fd::THREADS, fd::SMEM_BYTES));
  if (per_sm < 1) {
    set_error("decode_fused: kernel does not fit one CTA per SM (%d threads, %zu B shared memory)", fd::THREADS, fd::SMEM_BYTES);
    return KW_ERR_UNSUPPORTED;
  }
  const int G = f->n_sm, d = c.d_model, F = c.ffn_dim, V = c.vocab_size, S = c.max_source_pos, L = c.dec_layers;
  fd::Params& p = f->p;
  memset(&p, 0, sizeof(p));
  if (!make_cfg(p.g_qkv, 3 * d, d, G, true) || !make_cfg(p.g_dd, d, d, G, true) || !make_cfg(p.g_fc1, F, d, G, false) ||
      !make_cfg(p.g_fc2, d, F, G, true)) {
    set_error("decode_fused: no projection tiling for d=%d ffn=%d on %d SMs", d, F, G);
    return KW_ERR_UNSUPPORTED;
  }
  {  // vocabulary: tiles of R rows walked cta, cta + G, ...; R chosen so that every CTA gets the same number of tiles
    fd::GemmCfg& g = p.g_voc;
    const int per_cta = (V + G * 128 - 1) / (G * 128);
    g.N = V; g.K = d; g.S = 1; g.n_groups = G;
    g.R = std::min(128, round_up((V + G * per_cta - 1) / (G * per_cta), 8));
    g.rpad = g.R; g.kbps = 1; g.nkb = d / fd::BK;
    g.tiles = (V + g.R - 1) / g.R;
  }
  p.L = L; p.d = d; p.H = c.n_heads; p.F = F; p.V = V; p.S = S; p.MT = c.max_target_pos; p.G = G;
  p.kv_sw128 = getenv("KW_FUSED_KVSW") ? atoi(getenv("KW_FUSED_KVSW")) : 0;
  p.dbg_skip = getenv("KW_FUSED_SKIP") ? atoi(getenv("KW_FUSED_SKIP")) : 0;
  p.TR = (S % 120 == 0) ? 120 : 128;  // 8-row multiple: the swizzle pattern of a tile then starts at row phase 0
  p.n_kv_tiles = (S + p.TR - 1) / p.TR;

  const size_t es = 2;
  const size_t self_stride = (size_t)c.max_batch * d * c.max_target_pos * es;
  const size_t xkv_stride = (size_t)c.max_batch * S * 2 * d * es;
  std::vector<CUtensorMap> maps((size_t)L * fd::MAPS_PER_LAYER + fd::MAPS_GLOBAL);
  std::vector<fd::LayerP> layers(L);
  for (int l = 0; l < L; ++l) {
    const kw_dec_layer_weights& w = m->dec[l];
    CUtensorMap* lm = maps.data() + (size_t)l * fd::MAPS_PER_LAYER;
    int rc = 0;
    rc |= map2d(lm + fd::M_WQKV, w.wqkv, 3 * d, d, d, p.g_qkv.R);
    rc |= map2d(lm + fd::M_WO, w.wo, d, d, d, p.g_dd.R);
    rc |= map2d(lm + fd::M_WQX, w.wq_x, d, d, d, p.g_dd.R);
    rc |= map2d(lm + fd::M_WOX, w.wo_x, d, d, d, p.g_dd.R);
    rc |= map2d(lm + fd::M_W1, w.w1, F, d, d, p.g_fc1.R);
    rc |= map2d(lm + fd::M_W2, w.w2, d, F, F, p.g_fc2.R);
    rc |= map2d(lm + fd::M_XKV, (const char*)m->xkv + l * xkv_stride, c.max_batch * S, 2 * d, 2 * d, p.TR, p.kv_sw128 != 0);
    if (rc) return KW_ERR_CUDA;
    fd::LayerP& lp = layers[l];
    lp.ln1_w = w.ln1_w; lp.ln1_b = w.ln1_b; lp.lnx_w = w.lnx_w; lp.lnx_b = w.lnx_b; lp.ln3_w = w.ln3_w; lp.ln3_b = w.ln3_b;
    lp.bqkv = w.bqkv; lp.bo = w.bo; lp.bq_x = w.bq_x; lp.bo_x = w.bo_x; lp.b1 = w.b1; lp.b2 = w.b2;
    lp.self_k = (bf16*)((char*)m->self_k + l * self_stride);
    lp.self_v = (bf16*)((char*)m->self_v + l * self_stride);
  }
  {
    CUtensorMap* gm = maps.data() + (size_t)L * fd::MAPS_PER_LAYER;
    int rc = 0;
    rc |= map2d(gm + fd::M_VOCAB, m->w.tok_embed, V, d, d, p.g_voc.R);
    rc |= map2d(gm + fd::M_DA, m->da, c.max_batch, d, d, fd::NB);
    rc |= map2d(gm + fd::M_DATTN, m->dattn, c.max_batch, d, d, fd::NB);
    rc |= map2d(gm + fd::M_DH, m->dh, c.max_batch, F, F, fd::NB);
    if (rc) return KW_ERR_CUDA;
  }
  KW_CUDA_OK(cudaMalloc(&f->d_maps, maps.size() * sizeof(CUtensorMap)));
  KW_CUDA_OK(cudaMemcpy(f->d_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  KW_CUDA_OK(cudaMalloc(&f->d_layers, layers.size() * sizeof(fd::LayerP)));
  KW_CUDA_OK(cudaMemcpy(f->d_layers, layers.data(), layers.size() * sizeof(fd::LayerP), cudaMemcpyHostToDevice));
  KW_CUDA_OK(cudaMalloc(&f->part, (size_t)fd::MAX_SPLIT * fd::NB * d * sizeof(float)));
  KW_CUDA_OK(cudaMalloc(&f->partq, (size_t)fd::MAX_SPLIT * fd::NB * d * sizeof(float)));
  KW_CUDA_OK(cudaMalloc(&f->partqkv, (size_t)fd::MAX_SPLIT * fd::NB * 3 * d * sizeof(float)));
  KW_CUDA_OK(cudaMalloc(&f->vpart, (size_t)fd::NB * p.g_voc.tiles * 4 * fd::VP_WORDS * sizeof(float)));
  KW_CUDA_OK(cudaMalloc(&f->bar_counter, 256));
  KW_CUDA_OK(cudaMalloc(&f->steps_dev, 256));
  if (getenv("KW_FUSED_DEBUG")) {
    KW_CUDA_OK(cudaHostAlloc(&f->dbg_host, ((size_t)G * 16 + 64) * sizeof(unsigned), cudaHostAllocMapped));
    memset(f->dbg_host, 0, ((size_t)G * 16 + 64) * sizeof(unsigned));
    unsigned* dptr = nullptr;
    KW_CUDA_OK(cudaHostGetDevicePointer(&dptr, f->dbg_host, 0));
    p.dbg = dptr;
  }
  p.maps = f->d_maps;
  p.layers = f->d_layers;
  p.x = m->dx; p.partqkv = f->partqkv; p.part = f->part; p.partq = f->partq; p.vpart = f->vpart;
  p.da = (bf16*)m->da; p.dattn = (bf16*)m->dattn; p.dh = (bf16*)m->dh;
  p.tok_embed = (const bf16*)m->w.tok_embed;
  p.dec_pos = m->w.dec_pos; p.lnf_w = m->w.dec_ln_w; p.lnf_b = m->w.dec_ln_b;
  p.flags = m->flags;
  p.rules = m->rules;
  p.finished = m->finished;
  p.bar_counter = f->bar_counter;
  p.steps_out = f->steps_dev;
  f->ok = true;
  return KW_OK;
}

// One greedy pass (prompt already in `tokens`, cross K/V already projected) on the fused kernel.  Synchronises the
// stream; returns the number of decoder positions evaluated or a negative status.
int fused_decode_pass(kw_model* m, int B, int n_prompt, int max_length, int return_ts, int* tokens, cudaStream_t st) {
  FusedDecode* f = m->fused;
  if (!f || !f->ok) return KW_ERR_UNSUPPORTED;
  fd::Params p = f->p;
  p.B = B;
  p.tokens = tokens;
  p.ld_tokens = max_length;
  p.n_prompt = n_prompt;
  p.max_length = max_length;
  p.return_ts = return_ts;
  KW_CUDA_OK(cudaMemsetAsync(f->bar_counter, 0, 4, st));
  void* args[] = {&p};
  KW_CUDA_OK(cudaLaunchCooperativeKernel((void*)fd::dec_fused_kernel, dim3(f->n_sm), dim3(fd::THREADS), args,
                                         fd::SMEM_BYTES, st));
  ++g_launches;
  int steps = 0;
  cudaError_t e = cudaMemcpyAsync(&steps, f->steps_dev, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    if (f->dbg_host) {
      fprintf(stderr, "kwb200 decode_fused failed (%s); breadcrumbs [cta: phase | wait sites...]:\n", cudaGetErrorString(e));
      for (int c = 0; c < f->n_sm; ++c) {
        const unsigned* d = f->dbg_host + c * 16;
        bool any = false;
        for (int i = 4; i < 10; ++i) any |= d[i] != 0;
        if (c < 4 || any)
          fprintf(stderr, "  cta %3d: phase %u | W %08x KV %08x A %08x MMA %08x compute %08x gridbar %08x\n", c, d[0], d[4], d[5], d[6], d[7], d[8], d[9]);
      }
    }
    set_error("decode_fused: %s", cudaGetErrorString(e));
    return KW_ERR_CUDA;
  }
  if (f->dbg_host && getenv("KW_FUSED_TIMING")) {
    static const char* names[11] = {"row (res+LN)", "qkv gemm", "self-attn", "o gemm", "qx gemm", "cross-attn", "ox gemm",
                                    "fc1 gemm", "fc2 gemm", "vocab gemm", "combine+embed"};
    const unsigned* t = f->dbg_host + f->n_sm * 16;
    fprintf(stderr, "decode_fused timing (CTA 0, %d positions, B=%d): phase | work us | barrier us | count | per phase us\n", steps, B);
    double tot = 0;
    for (int i = 0; i < 11; ++i) {
      fprintf(stderr, "  %-14s %8u %8u %6u   %6.2f + %5.2f\n", names[i], t[i * 3], t[i * 3 + 1], t[i * 3 + 2],
              t[i * 3 + 2] ? (double)t[i * 3] / t[i * 3 + 2] : 0.0, t[i * 3 + 2] ? (double)t[i * 3 + 1] / t[i * 3 + 2] : 0.0);
      tot += t[i * 3] + t[i * 3 + 1];
    }
    fprintf(stderr, "  total %.1f us = %.1f us per position\n", tot, steps ? tot / steps : 0.0);
  }
  return steps;
}

}  // namespace kw
