// tcgen05 flash attention for the Whisper encoder (non-causal, no mask, head dim 64, q pre-scaled; the arithmetic of
// sdpa_attention_forward as called at modeling_whisper.py:342-352), bf16 operands, fp32 softmax / accumulation.
//
// CTA = one 128-query tile of one (batch, head) walking 64-key tiles.  Everything but Q / K / V lives in TENSOR MEMORY:
//   warp 4, lane 0 issues both the loads and the MMAs:
//            TMA:  Q once, then K and V tiles ([64 keys x 64] bf16, 128 B swizzle, 3-slot rings) via 3D tensor maps over the
//                  strided q/k/v views (coordinates = column, time, batch; out-of-range rows are zero-filled)
//            MMA:  S(j) = Q K_j^T   tcgen05.mma M=128 N=64 K=64 (both K-major)            -> TMEM S  (64 columns)
//                  O += P(j) V_j    tcgen05.mma M=128 N=64 K=64, A = P straight from TMEM, V MN-major from smem -> TMEM O
//   warps 0-3      ONE thread per query row: tcgen05.ld of the row's 64 scores, S handed back to the MMA warp at once
//                  (S(j+1) is computed while this tile's arithmetic runs: the scores in flight live in registers, not
//                  in a second TMEM buffer), row max / lazy running max, p = exp2(s log2e - m log2e) with packed
//                  fp32x2 FMA / ADD and MUFU.EX2, row sum, P -> packed bf16 pairs into its own 32 TMEM columns
//                  (tcgen05.st), O rescaled in TMEM only when the running max moved by more than 2^8.
// 160 TMEM columns per CTA from TWO allocations (128: S and O; 32: P) and <= 128 registers put THREE CTAs on an SM; their
// phases interleave on the MUFU, which ends up ~90 % busy (the practical bound at head dim 64: 13.5 exp2 / clk / SM).
// Measured history of this kernel (B=64, H=20, T=1500, isolated): 2 CTAs/SM with S and P double buffered in 256 columns
// 638 TFLOP/s; 4 CTAs/SM with P written over S in 128 columns (serial S -> softmax -> PV chain per CTA) 672; this
// layout 767.  Finally O / l is written as bf16.
#include <atomic>

#include "tc_common.cuh"

namespace kw {

extern std::atomic<long long> g_launches;

namespace tc {

constexpr int ABQ = 128, ABK = 64, AHD = 64;
constexpr int TILE_BYTES = 128 * 64 * 2;           // Q tile: 16 KB
constexpr int KV_BYTES = ABK * 64 * 2;             // K, V tiles: 8 KB each
constexpr int ATT_SM_WARPS = 4;                    // softmax warps: one thread per query row (all 64 keys of a tile)
constexpr int ATT_SM_THREADS = ATT_SM_WARPS * 32;
constexpr int ATT_THREADS = 32 * (ATT_SM_WARPS + 1);   // + one warp whose lane 0 issues both the TMA loads and the MMAs
constexpr uint32_t IDESC_S = make_idesc(128, ABK, 0, 0);
constexpr uint32_t IDESC_PV = make_idesc(128, 64, 0, 1);  // B = V is MN-major (head dim contiguous)

struct AttnParams {
  bf16* out;
  long long o_sb, o_st;
  int Tq, Tk;
  int v_lbo, v_sbo, v_kstep;  // V (MN-major) descriptor parameters, bytes
};

#ifdef KW_ATT_TIMING
// debug build only: per-phase cycle sums of the softmax loop (lane 0 of warp 0 of every CTA)
__device__ unsigned long long g_att_dbg[8];
#define ATT_T(i)                                                      \
  do {                                                                \
    if (threadIdx.x == 0) {                                           \
      const long long now_ = clock64();                               \
      dbg_acc[i] += now_ - dbg_t;                                     \
      dbg_t = now_;                                                   \
    }                                                                 \
  } while (0)
#else
#define ATT_T(i)
#endif
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for x <= 0 on the FMA / ALU pipes instead of the 16-lane MUFU: round-to-nearest split x = n + f, |f| <= 0.5, a
// degree-3 minimax polynomial for 2^f (max relative error 7.5e-5, 50x below the bf16 rounding of P) and n added into the
// exponent field.  Inputs below -125 (masked keys: -inf) return ~2^-125 instead of 0; those keys meet all-zero V rows.
// Measured on B200 (B=64, H=20, T=1500): never a win — 590 / 564 TFLOP/s against 609 at 25 % / 50 % offload on the earlier
// 2-CTA kernel (latency-chain bound), 743 / 652 against 767 on this kernel (111 / 128 registers, no spills): a warp's own
// instruction stream, not the shared MUFU, sets the length of its exp phase.  The packed fp32x2 form below (KW_ATT_POLY_PACKED,
// 118 registers) brings the 25 % offload level with the plain kernel, 763.6 against 766.2 TFLOP/s, and no further
// (profiles/r2_attention_poly_variants.txt).  Compiled out by default.
#ifndef KW_ATT_POLY_PAIRS
#define KW_ATT_POLY_PAIRS 0  // pairs out of every 4 (8 scores) whose exp2 runs on the FMA pipe
#endif
__device__ __forceinline__ float ex2_poly(float x) {  // scalar form: every constant is an instruction immediate
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;  // 1.5 * 2^23: mantissa low bits = round(x)
  const float f = x - (t - 12582912.0f);
  float q = fmaf(0.05517125502228737f, f, 0.2426103800535202f);
  q = fmaf(q, f, 0.6932609677314758f);
  q = fmaf(q, f, 0.9999281167984009f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));
}
#ifndef KW_ATT_POLY_PACKED
#define KW_ATT_POLY_PACKED 0
#endif
#if KW_ATT_POLY_PACKED
// the same arithmetic on the packed fp32x2 pipe: 2 FMNMX + 2 FADD2 + 4 FFMA2 + 2 (shift + add) per PAIR instead of 9 scalar
// operations per element
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x = make_float2(fmaxf(x.x, -125.0f), fmaxf(x.y, -125.0f));
  const float2 M = make_float2(12582912.0f, 12582912.0f), NM = make_float2(-12582912.0f, -12582912.0f);
  const float2 t = __fadd2_rn(x, M);
  const float2 f = __ffma2_rn(__fadd2_rn(t, NM), make_float2(-1.0f, -1.0f), x);
  float2 q = __ffma2_rn(make_float2(0.05517125502228737f, 0.05517125502228737f), f, make_float2(0.2426103800535202f, 0.2426103800535202f));
  q = __ffma2_rn(q, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  q = __ffma2_rn(q, f, make_float2(0.9999281167984009f, 0.9999281167984009f));
  return make_float2(__int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23)));
}
#else
__device__ __forceinline__ float2 ex2_poly2(float2 x) { return make_float2(ex2_poly(x.x), ex2_poly(x.y)); }
#endif
// Measured and dropped in round 2: skipping the exp2 work of query rows past Tq (a warp of the last 128-row tile) and of
// the key groups past Tk in the last key tile (4.4 % of the MUFU work in total): 760 / 666 TFLOP/s against 767 — the
// extra control flow costs registers (109 / 128 against 106) and scheduling freedom in the loop that matters.
// (Packed half-precision exponentials — ex2.approx.f16x2 / ex2.approx.ftz.bf16x2 — were checked as a way to halve the
// MUFU work: ptxas lowers both to TWO scalar MUFU.EX2.F16 / .BF16 operations on sm_100a, so there is nothing to gain.)

#ifndef KW_ATT3_KVS
#define KW_ATT3_KVS 3
#endif
namespace a3 {
constexpr int KVS = KW_ATT3_KVS;
constexpr int OFF_Q = 0, OFF_K = TILE_BYTES, OFF_V = OFF_K + KVS * KV_BYTES, OFF_BAR = OFF_V + KVS * KV_BYTES;
// Padded to > 227 KB / 4 so that shared memory, not the register count of some future edit, caps residency at three CTAs
// per SM: a fourth CTA would take the last 128 TMEM columns and all four would wait forever for their 32-column block.
constexpr size_t SMEM_USED = 1024 + OFF_BAR + 256;
constexpr size_t SMEM = SMEM_USED > 58 * 1024 ? SMEM_USED : 58 * 1024;
static_assert(4 * SMEM > 227 * 1024 && 3 * SMEM <= 227 * 1024, "attn_tc3_kernel must fit exactly three CTAs per SM");
}  // namespace a3

__global__ void __launch_bounds__(ATT_THREADS, 3)
attn_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using namespace a3;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + a3::OFF_BAR;
  const uint32_t q_full = bar0;
  auto k_full = [&](int s) { return bar0 + 8u * (1 + s); };
  auto v_full = [&](int s) { return bar0 + 8u * (1 + KVS + s); };
  constexpr int BS = 1 + 2 * KVS;
  const uint32_t s_full = bar0 + 8u * BS, s_empty = bar0 + 8u * (BS + 1), p_full = bar0 + 8u * (BS + 2),
                 o_full = bar0 + 8u * (BS + 3);
  const uint32_t tmem_slot = bar0 + 8u * (BS + 4);  // two 32-bit slots: the 128-column block, the 32-column block
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gen_base + a3::OFF_BAR + 8 * (BS + 4));
  static_assert(8 * (BS + 5) <= 256, "barrier area");
  static_assert(KVS == 3, "the slot-reuse argument of the issuing thread below is written for a 3-slot ring");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ABQ, h = blockIdx.y, b = blockIdx.z;
  const int n_kt = (p.Tk + ABK - 1) / ABK;
  constexpr int W_MMA = ATT_SM_WARPS;

  if (warp == W_MMA && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    mbar_init(q_full, 1);
    for (int s = 0; s < KVS; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(v_full(s), 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, ATT_SM_THREADS);
    mbar_init(p_full, ATT_SM_THREADS);
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {  // both allocations, then the permit is given up (three resident CTAs: 480 of 512 columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(128) : "memory");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot + 4), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_a = tmem_slot_ptr[0], tmem_b = tmem_slot_ptr[1];
  const uint32_t tS = tmem_a, tO = tmem_a + 64, tP = tmem_b;

  if (warp == W_MMA) {
    // ===================== TMA + MMA issue: one thread =====================
    // No producer warp and no "slot empty" barriers: by the time this thread has seen p_full(j), S(j) has retired (the
    // softmax threads read it), so K slot j % 3 is free for K(j+3); and every softmax thread waited for P(j-1) V(j-1)
    // before it arrived on p_full(j), so V slot (j-1) % 3 is free for V(j+2).  Five warps per CTA also make 15 warps
    // per SM, which divide over the four schedulers without the fifth warp that capped the kernel at 96 registers.
    if (lane == 0) {
      auto load_k = [&](int j) {
        const int s = j % KVS;
        mbar_expect_tx(k_full(s), KV_BYTES);
        tma_load_3d(base + a3::OFF_K + s * KV_BYTES, &tmK, k_full(s), h * AHD, j * ABK, b);
      };
      auto load_v = [&](int j) {
        const int s = j % KVS;
        mbar_expect_tx(v_full(s), KV_BYTES);
        tma_load_3d(base + a3::OFF_V + s * KV_BYTES, &tmV, v_full(s), h * AHD, j * ABK, b);
      };
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_3d(base + a3::OFF_Q, &tmQ, q_full, h * AHD, q0, b);
      for (int j = 0; j < KVS && j < n_kt; ++j) {
        load_k(j);
        load_v(j);
      }
      mbar_wait(q_full, 0);
      const uint64_t dq = make_desc(base + a3::OFF_Q);
      auto issue_S = [&](int j) {
        const int ks = j % KVS;
        mbar_wait(k_full(ks), (j / KVS) & 1);
        tc_fence_after();
        const uint64_t dk = make_desc(base + a3::OFF_K + ks * KV_BYTES);
#pragma unroll
        for (int k = 0; k < AHD / 16; ++k) umma_f16(tS, dq + 2 * k, dk + 2 * k, IDESC_S, k != 0);
        umma_commit(s_full);
      };
      issue_S(0);
      for (int j = 0; j < n_kt; ++j) {
        const int vs = j % KVS;
        if (j + 1 < n_kt) {
          mbar_wait(s_empty, j & 1);  // every row's S(j) is in registers
          issue_S(j + 1);
        }
        mbar_wait(p_full, j & 1);  // P(j) written, O rescaled
        mbar_wait(v_full(vs), (j / KVS) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < ABK / 16; ++k) {
          const uint64_t dv = make_desc_sw128(base + a3::OFF_V + vs * KV_BYTES + k * p.v_kstep, p.v_lbo, p.v_sbo);
          umma_f16_ts(tO, tP + k * 8, dv, IDESC_PV, (j | k) != 0);  // 16 keys = 8 packed columns
        }
        umma_commit(o_full);
        if (j + KVS < n_kt) load_k(j + KVS);
        if (j >= 1 && j + 2 < n_kt) load_v(j + 2);
      }
    }
  } else {
    const int r = warp * 32 + lane;                      // query row of the tile = TMEM lane
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    // Running maximum in the log2 domain (mb = m * log2 e).  It is only moved when a tile's maximum exceeds it by more
    // than RESCALE_T (p <= 2^8 in between: harmless for the fp32 sums and for bf16 P, and the common offset cancels in
    // O / l exactly as with the tight maximum), so the TMEM round trip that rescales O happens a few times per row block
    // instead of on nearly every tile.
    constexpr float RESCALE_T = 8.0f;
    float mb_run = -INFINITY, l_run = 0.0f;
    const float2 L2 = make_float2(LOG2E, LOG2E);
#ifdef KW_ATT_TIMING
    long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dbg_t = clock64();
#endif
    for (int j = 0; j < n_kt; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      ATT_T(0);
      uint32_t v[ABK];
      tmem_ld32(tS + lane_off, v);
      tmem_ld32(tS + lane_off + 32, v + 32);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      mbar_arrive(s_empty);                               // S(j+1) may overwrite the score columns now
      ATT_T(1);
      const int kbase = j * ABK;
      if (kbase + ABK > p.Tk) {
#pragma unroll
        for (int i = 0; i < ABK; ++i)
          if (kbase + i >= p.Tk) v[i] = 0xff800000u;      // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < ABK; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(v[i]));
        mx1 = fmaxf(mx1, __uint_as_float(v[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(v[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(v[i + 3]));
      }
      const float mt = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * LOG2E;  // finite: every tile holds >= 1 valid key
      const bool moved = mt > mb_run + RESCALE_T;           // always on the first tile (mb_run = -inf)
      const float mb = moved ? mt : mb_run;
      const float alpha = moved ? ex2(mb_run - mb) : 1.0f;  // 0 on the first tile
      const float2 nmb = make_float2(-mb, -mb);
      ATT_T(2);
      float2 rs0 = make_float2(0.0f, 0.0f), rs1 = make_float2(0.0f, 0.0f);
      uint32_t pk[ABK / 2];
#pragma unroll
      for (int i = 0; i < ABK; i += 8) {
        float2 e[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[i + 2 * q]), __uint_as_float(v[i + 2 * q + 1])), L2, nmb);
          e[q] = q >= 4 - KW_ATT_POLY_PAIRS ? ex2_poly2(x) : make_float2(ex2(x.x), ex2(x.y));  // exp2(-inf) = 0: masked keys
          pk[(i >> 1) + q] = pack_bf16(e[q].x, e[q].y);
        }
        rs0 = __fadd2_rn(rs0, __fadd2_rn(e[0], e[2]));
        rs1 = __fadd2_rn(rs1, __fadd2_rn(e[1], e[3]));
      }
      l_run = l_run * alpha + ((rs0.x + rs0.y) + (rs1.x + rs1.y));
      mb_run = mb;
      ATT_T(3);
      if (j > 0) {  // P(j-1) V(j-1) complete: the P columns are free again and O is stable
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
      }
      ATT_T(4);
      tmem_st32(tP + lane_off, pk);
      if (j > 0 && __any_sync(0xffffffffu, moved)) {
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          uint32_t o[32];
          tmem_ld32(tO + lane_off + h2 * 32, o);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tO + lane_off + h2 * 32, o);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");  // P (and the rescaled O) are in TMEM
      tc_fence_before();
      mbar_arrive(p_full);
      ATT_T(5);
    }
    mbar_wait(o_full, (n_kt - 1) & 1);
    tc_fence_after();
    ATT_T(6);
#ifdef KW_ATT_TIMING
    if (threadIdx.x == 0) {
      for (int i = 0; i < 7; ++i) atomicAdd(&g_att_dbg[i], (unsigned long long)dbg_acc[i]);
      atomicAdd(&g_att_dbg[7], 1ull);
    }
#endif
    const int t = q0 + r;
    const float inv = 1.0f / l_run;
    bf16* orow = p.out + (size_t)b * p.o_sb + (size_t)t * p.o_st + h * AHD;
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      uint32_t o[32];
      tmem_ld32(tO + lane_off + h2 * 32, o);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (t < p.Tq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          w.y = pack_bf16(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          w.z = pack_bf16(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          w.w = pack_bf16(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + h2 * 32 + i) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tmem_dealloc(tmem_a, 128);
    tmem_dealloc(tmem_b, 32);
  }
}

static int make_map3(CUtensorMap* map, const void* ptr, int B, int T, int H, long long sb, long long st, int box_rows) {
  cuuint64_t gdim[3] = {(cuuint64_t)H * AHD, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t gstride[2] = {(cuuint64_t)st * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[3] = {(cuuint32_t)AHD, (cuuint32_t)box_rows, 1};
  return make_map_bf16(map, ptr, 3, gdim, gstride, box);
}

static int g_v_lbo = 1024, g_v_sbo = 1024, g_v_kstep = 2048;  // N = 64 is one MN block: only the 8-row K-group stride matters

}  // namespace tc

void attention_tc_debug(int lbo, int sbo, int kstep) {
#ifdef KW_ATT_TIMING
  if (lbo < 0) {  // dump + reset the phase counters
    unsigned long long h[8], z[8] = {0};
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, tc::g_att_dbg, sizeof(h));
    const char* nm[7] = {"0 wait S(j)", "1 tcgen05.ld scores + release S", "2 mask + max + alpha", "3 ffma / ex2 / pack / sum",
                         "4 wait PV(j-1)", "5 st P (+ rare O rescale) + arrive", "6 final wait for O"};
    const double ctas = (double)h[7];
    for (int i = 0; i < 7; ++i) fprintf(stderr, "  att phase %-32s %10.1f cycles per CTA\n", nm[i], h[i] / ctas);
    cudaMemcpyToSymbol(tc::g_att_dbg, z, sizeof(z));
    return;
  }
#endif
  tc::g_v_lbo = lbo;
  tc::g_v_sbo = sbo;
  tc::g_v_kstep = kstep;
}

int attention_tc(const void* q, const void* k, const void* v, void* out, int B, int H, int Tq, int Tk, long long q_sb,
                 long long q_st, long long kv_sb, long long kv_st, long long o_sb, long long o_st, cudaStream_t st) {
  using namespace tc;
  if (B < 1 || H < 1 || Tq < 1 || Tk < 1 || H > 65535 || B > 65535) return KW_ERR_UNSUPPORTED;
  if ((q_st % 8) || (kv_st % 8) || (q_sb % 8) || (kv_sb % 8) || (o_st % 8) || (o_sb % 8)) return KW_ERR_UNSUPPORTED;
  if (((uintptr_t)q & 15) || ((uintptr_t)k & 15) || ((uintptr_t)v & 15) || ((uintptr_t)out & 15)) return KW_ERR_UNSUPPORTED;
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_map3(&tmQ, q, B, Tq, H, q_sb, q_st, ABQ);
  if (rc) return rc;
  if ((rc = make_map3(&tmK, k, B, Tk, H, kv_sb, kv_st, ABK))) return rc;
  if ((rc = make_map3(&tmV, v, B, Tk, H, kv_sb, kv_st, ABK))) return rc;
  static bool attr = false;
  if (!attr) {
    KW_CUDA_OK(cudaFuncSetAttribute(attn_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a3::SMEM));
    attr = true;
  }
  AttnParams p;
  p.out = (bf16*)out; p.o_sb = o_sb; p.o_st = o_st; p.Tq = Tq; p.Tk = Tk;
  p.v_lbo = g_v_lbo; p.v_sbo = g_v_sbo; p.v_kstep = g_v_kstep;
  dim3 grid(ceil_div(Tq, ABQ), H, B);
  attn_tc3_kernel<<<grid, ATT_THREADS, a3::SMEM, st>>>(tmQ, tmK, tmV, p);
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

}  // namespace kw
