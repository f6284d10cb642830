// tcgen05 flash attention for the Whisper encoder (non-causal, no mask, head dim 64, q pre-scaled; the arithmetic of
// sdpa_attention_forward as called at modeling_whisper.py:342-352), bf16 operands, fp32 softmax / accumulation.
//
// CTA = one 128-query tile of one (batch, head); two CTAs are resident per SM (80 KB smem, 256 TMEM columns each) so
// one CTA's softmax overlaps the other's MMAs.  Per 128-key tile:
//   warp 4  TMA:  K tile and V tile ([128 keys x 64] bf16, 128 B swizzle) via 3D tensor maps over the strided q/k/v views
//                 (coordinates = column, time, batch; out-of-range rows are zero-filled)
//   warp 5  MMA:  S = Q K^T        tcgen05.mma M=128 N=128 K=64 (Q, K both K-major)      -> TMEM columns [0,128)
//                 O += P V         tcgen05.mma M=128 N=64  K=128 (P K-major from smem, V MN-major) -> TMEM columns [128,192)
//   warps 0-7     two threads per query row (64 keys each; warp w owns TMEM lanes 32*(w%4).. and key half w/4): partial
//                 row max over S (tcgen05.ld) exchanged through smem, rescale of O in TMEM when the running max moves
//                 (tcgen05.ld / tcgen05.st, 32 columns per thread), p = exp2(s*log2e - m*log2e), partial row sums,
//                 P -> bf16 into swizzled smem (one 64-key swizzle atom per thread half).
// Finally O / l is written as bf16, 64 bytes per thread.
#include <atomic>

#include "tc_common.cuh"

namespace kw {

extern std::atomic<long long> g_launches;

namespace tc {

constexpr int ABQ = 128, ABK = 128, AHD = 64;
constexpr int TILE_BYTES = 128 * 64 * 2;           // Q, K, V tiles: 16 KB each
constexpr int P_BYTES = 2 * TILE_BYTES;            // P: two [128 x 64-key] swizzle atoms
constexpr int ATT_SM_WARPS = 8;                     // softmax warps: 2 threads per query row (64 keys each)
constexpr int ATT_THREADS = 32 * (ATT_SM_WARPS + 2);
constexpr int ATT_TMEM_COLS = 256;                 // S: 128 columns, O: 64 columns
constexpr size_t ATT_SMEM = 1024 + 3 * TILE_BYTES + P_BYTES + 128 + 2 * 128 * 4;
constexpr uint32_t IDESC_S = make_idesc(128, 128, 0, 0);
constexpr uint32_t IDESC_PV = make_idesc(128, 64, 0, 1);  // B = V is MN-major (head dim contiguous)

struct AttnParams {
  bf16* out;
  long long o_sb, o_st;
  int Tq, Tk;
  int v_lbo, v_sbo, v_kstep;  // V (MN-major) descriptor parameters, bytes
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = base + TILE_BYTES, sV = base + 2 * TILE_BYTES, sP = base + 3 * TILE_BYTES;
  const uint32_t bar0 = sP + P_BYTES;
  const uint32_t q_full = bar0, k_full = bar0 + 8, k_empty = bar0 + 16, v_full = bar0 + 24, v_empty = bar0 + 32,
                 s_full = bar0 + 40, p_full = bar0 + 48, o_full = bar0 + 56, tmem_slot = bar0 + 64;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gen_base + 3 * TILE_BYTES + P_BYTES + 64);
  uint8_t* P_gen = gen_base + 3 * TILE_BYTES;
  float* s_xchg = reinterpret_cast<float*>(gen_base + 3 * TILE_BYTES + P_BYTES + 128);  // [2][128] partial row max / partial row sum

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ABQ, h = blockIdx.y, b = blockIdx.z;
  const int n_kt = (p.Tk + ABK - 1) / ABK;

  if (warp == ATT_SM_WARPS && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_full, ATT_SM_WARPS * 32); mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == ATT_SM_WARPS + 1) tmem_alloc(tmem_slot, ATT_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == ATT_SM_WARPS) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_3d(sQ, &tmQ, q_full, h * AHD, q0, b);
      for (int j = 0; j < n_kt; ++j) {
        if (j > 0) mbar_wait(k_empty, (j - 1) & 1);
        mbar_expect_tx(k_full, TILE_BYTES);
        tma_load_3d(sK, &tmK, k_full, h * AHD, j * ABK, b);
        if (j > 0) mbar_wait(v_empty, (j - 1) & 1);
        mbar_expect_tx(v_full, TILE_BYTES);
        tma_load_3d(sV, &tmV, v_full, h * AHD, j * ABK, b);
      }
    }
  } else if (warp == ATT_SM_WARPS + 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      mbar_wait(q_full, 0);
      const uint64_t dq = make_desc(sQ), dk = make_desc(sK);
      auto issue_S = [&](int j) {  // S_j = Q K_j^T
        mbar_wait(k_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AHD / 16; ++k) umma_f16(tS, dq + 2 * k, dk + 2 * k, IDESC_S, k != 0);
        umma_commit(k_empty);
        umma_commit(s_full);
      };
      issue_S(0);
      for (int j = 0; j < n_kt; ++j) {
        // p_full(j): the softmax threads have finished reading S_j and written P_j
        mbar_wait(p_full, j & 1);
        // S_{j+1} goes first so the next softmax can start while P_j V_j is still running on the tensor pipe
        if (j + 1 < n_kt) issue_S(j + 1);
        mbar_wait(v_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < ABK / 16; ++k) {
          const uint64_t dp = make_desc(sP + (k / 4) * TILE_BYTES) + 2 * (k % 4);
          const uint64_t dv = make_desc_sw128(sV + k * p.v_kstep, p.v_lbo, p.v_sbo);
          umma_f16(tO, dp, dv, IDESC_PV, (j | k) != 0);
        }
        umma_commit(v_empty);
        umma_commit(o_full);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue: two threads per query row =====================
    const int quarter = warp & 3, hc = warp >> 2;        // TMEM lane quarter, key half
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    float m_run = -INFINITY, l_run = 0.0f;
    for (int j = 0; j < n_kt; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * ABK + hc * 64;                // first key of this thread's half
      const bool tail = j * ABK + ABK > p.Tk;
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tS + lane_off + hc * 64 + c * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kbase + c * 32 + i >= p.Tk) v[i] = 0xff800000u;  // -inf
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(v[i]));
          mx1 = fmaxf(mx1, __uint_as_float(v[i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(v[i + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(v[i + 3]));
        }
      }
      const float pm = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      s_xchg[hc * 128 + r] = pm;
      asm volatile("bar.sync 1, 256;" ::: "memory");      // the 256 softmax threads only
      const float m_new = fmaxf(m_run, fmaxf(pm, s_xchg[(hc ^ 1) * 128 + r]));  // finite: >= 1 valid key per tile
      const float alpha = ex2((m_run - m_new) * LOG2E);   // 0 on the first tile (m_run = -inf)
      const float mb = m_new * LOG2E;
      if (j > 0) {
        // PV_{j-1} has completed: O may be rescaled and P's smem may be overwritten
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
          uint32_t v[32];
          tmem_ld32(tO + lane_off + hc * 32, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st32(tO + lane_off + hc * 32, v);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
      }
      float rs0 = 0.0f, rs1 = 0.0f, rs2 = 0.0f, rs3 = 0.0f;
      uint8_t* prow = P_gen + hc * TILE_BYTES + r * 128;   // atom hc holds keys [64 hc, 64 hc + 64) of the tile
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tS + lane_off + hc * 64 + c * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float p0 = ex2(fmaf(__uint_as_float(v[i]), LOG2E, -mb));
          float p1 = ex2(fmaf(__uint_as_float(v[i + 1]), LOG2E, -mb));
          float p2 = ex2(fmaf(__uint_as_float(v[i + 2]), LOG2E, -mb));
          float p3 = ex2(fmaf(__uint_as_float(v[i + 3]), LOG2E, -mb));
          if (tail) {
            if (kbase + c * 32 + i >= p.Tk) p0 = 0.0f;
            if (kbase + c * 32 + i + 1 >= p.Tk) p1 = 0.0f;
            if (kbase + c * 32 + i + 2 >= p.Tk) p2 = 0.0f;
            if (kbase + c * 32 + i + 3 >= p.Tk) p3 = 0.0f;
          }
          rs0 += p0; rs1 += p1; rs2 += p2; rs3 += p3;
          pk[i >> 1] = pack_bf16(p0, p1);
          pk[(i >> 1) + 1] = pack_bf16(p2, p3);
        }
        // 16-byte chunks c*4 + {0..3} of this row's 128-byte atom line, XOR-swizzled with r % 8
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (c * 4 + q) ^ (r & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      l_run = l_run * alpha + ((rs0 + rs1) + (rs2 + rs3));
      m_run = m_new;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy P stores -> visible to the UMMA (async proxy)
      tc_fence_before();
      mbar_arrive(p_full);
    }
    mbar_wait(o_full, (n_kt - 1) & 1);
    tc_fence_after();
    // total row sum = sum of the two halves' partial sums (same running max on both sides).  Reusing the max-exchange
    // slots is safe: every thread read them before arriving on p_full(last), which PV_last / o_full(last) waited for.
    s_xchg[hc * 128 + r] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float l_all = l_run + s_xchg[(hc ^ 1) * 128 + r];
    const int t = q0 + r;
    const float inv = 1.0f / l_all;
    bf16* orow = p.out + (size_t)b * p.o_sb + (size_t)t * p.o_st + h * AHD + hc * 32;
    {
      uint32_t v[32];
      tmem_ld32(tO + lane_off + hc * 32, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (t < p.Tq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
          w.y = pack_bf16(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
          w.z = pack_bf16(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
          w.w = pack_bf16(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + i) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == ATT_SM_WARPS + 1) tmem_dealloc(tmem_base, ATT_TMEM_COLS);
}

static int make_map3(CUtensorMap* map, const void* ptr, int B, int T, int H, long long sb, long long st) {
  cuuint64_t gdim[3] = {(cuuint64_t)H * AHD, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t gstride[2] = {(cuuint64_t)st * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[3] = {(cuuint32_t)AHD, 128, 1};
  return make_map_bf16(map, ptr, 3, gdim, gstride, box);
}

static int g_v_lbo = 1024, g_v_sbo = 1024, g_v_kstep = 2048;  // N = 64 is one MN block: only the 8-row K-group stride matters

}  // namespace tc

void attention_tc_debug(int lbo, int sbo, int kstep) {
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
  int rc = make_map3(&tmQ, q, B, Tq, H, q_sb, q_st);
  if (rc) return rc;
  if ((rc = make_map3(&tmK, k, B, Tk, H, kv_sb, kv_st))) return rc;
  if ((rc = make_map3(&tmV, v, B, Tk, H, kv_sb, kv_st))) return rc;
  static bool attr = false;
  if (!attr) {
    KW_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM));
    attr = true;
  }
  AttnParams p;
  p.out = (bf16*)out; p.o_sb = o_sb; p.o_st = o_st; p.Tq = Tq; p.Tk = Tk;
  p.v_lbo = g_v_lbo; p.v_sbo = g_v_sbo; p.v_kstep = g_v_kstep;
  dim3 grid(ceil_div(Tq, ABQ), H, B);
  attn_tc_kernel<<<grid, ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, p);
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

}  // namespace kw
