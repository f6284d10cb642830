// tcgen05 / TMEM / TMA GEMM for sm_100a:  out[M,N] = epi(A[M,K] . W[N,K]^T + bias), bf16 operands, fp32 accumulation.
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 4      TMA producer: cp.async.bulk.tensor 2D loads of a 128 x 64 A tile and a 256 x 64 W tile (both K-major,
//               128-byte swizzle) into a 4-stage shared-memory ring, completion on mbarriers (complete_tx)
//   warp 5      MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16) x 4 per stage,
//               accumulating in TMEM; tcgen05.commit releases the smem stage and, after the last k-block, publishes the
//               accumulator.  Two 256-column accumulators (all 512 TMEM columns) let tile i+1's MMAs overlap tile i's epilogue
//   warps 0-3   epilogue: tcgen05.ld 32 lanes x 32 columns at a time -> bias / GELU(erf) / residual / position add in
//               registers -> direct 16-byte global stores (each thread owns one output row)
// Tiles are walked n-fastest so the CTAs running together share A row-blocks and the whole weight matrix stays in L2.
// Every mbarrier wait carries a clock64 watchdog that traps instead of hanging the GPU if a pipeline bug deadlocks.
#include <atomic>

#include "sample_rules.cuh"
#include "tc_common.cuh"

namespace kw {

extern std::atomic<long long> g_launches;

namespace tc {

#ifndef KW_EPI_WARPS
#define KW_EPI_WARPS 8
#endif
constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int N_EPI_WARPS = KW_EPI_WARPS, THREADS = 32 * (N_EPI_WARPS + 2);  // wide kernels: 16 epilogue warps + TMA + MMA
constexpr int EPI_THREADS = N_EPI_WARPS * 32, EPI_COLS = BN / (N_EPI_WARPS / 4);  // 64 columns per epilogue warp
constexpr int SK_EPI_WARPS = 4, SK_THREADS = 32 * (SK_EPI_WARPS + 2);  // skinny kernel
constexpr int TMEM_COLS = 512;
// per-warp staging block of the residual epilogue: 32 rows x 16 cols, either dense with the tensor map's 64 B swizzle
// (TMA reduction: 2 KB, 512 B aligned) or padded to 20 floats per row (read-modify-write path)
constexpr int EPI_STAGE_LD = 20, EPI_STAGE_FLOATS = 32 * EPI_STAGE_LD;
constexpr int EPI_STAGE_OFF = 256 /*barriers*/ + 2 * BN * 4 /*bias*/ + 256 /*pad: staging blocks 512 B aligned*/;
static_assert(EPI_STAGE_OFF % 512 == 0 && (EPI_STAGE_FLOATS * 4) % 512 == 0, "TMA reduce staging alignment");
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + (size_t)STAGES * STAGE_BYTES + EPI_STAGE_OFF +
                              KW_EPI_WARPS * EPI_STAGE_FLOATS * 4 /*residual epilogue staging*/;

struct Params {
  unsigned long long* stamps;  // debug timeline (globaltimer ns) written by CTA 0, or nullptr
  const float* bias;
  void* out;
  const float* pos;
  int M, N, K, ldo, pos_period, epi, out_bf16;
  SampleFuse sf;  // EPI_ARGMAX (decode-time vocabulary projection)
  int w_hint;  // skinny kernel: L2 eviction priority of the weight loads (GemmArgs::w_hint)
  int store_tma;  // wide kernels, bf16 outputs: 32 x 32 blocks leave through TMA stores instead of per-thread 16-byte stores
  int resid_tma;  // wide kernels, EPI_RESID: 1 = x += tile as a TMA reduction at the L2, 0 = read-modify-write by the epilogue warps
};

constexpr uint32_t IDESC = make_idesc(BM, BN, 0, 0);

// GELU(erf) for bf16 outputs: erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below bf16 rounding) with one
// MUFU.RCP and one MUFU.EX2 instead of libdevice erff's longer polynomial; the fp32 path keeps erff.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));  // 1 MUFU (the IEEE __frcp_rn is ~10 instr)
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
  const float erf_abs = fmaf(-poly, e, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// Two GELUs at once on the packed fp32x2 pipe (sm_100 FFMA2 / FMUL2): 12 packed ops + 2 LOP + 4 MUFU per pair instead of
// ~15 scalar ops + 2 MUFU per element.  Same A&S 7.1.26 arithmetic; gelu(x) = x/2 + |x|/2 * erf(|x|/sqrt 2).
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 den = __ffma2_rn(ax, make_float2(0.3275911f * 0.70710678118654752440f, 0.3275911f * 0.70710678118654752440f),
                                make_float2(1.0f, 1.0f));
  float2 t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
  // -poly(t) * t, coefficients negated so that erf = 1 + npoly * e is a single FMA
  float2 np = __ffma2_rn(t, make_float2(-1.061405429f, -1.061405429f), make_float2(1.453152027f, 1.453152027f));
  np = __ffma2_rn(np, t, make_float2(-1.421413741f, -1.421413741f));
  np = __ffma2_rn(np, t, make_float2(0.284496736f, 0.284496736f));
  np = __ffma2_rn(np, t, make_float2(-0.254829592f, -0.254829592f));
  np = __fmul2_rn(np, t);
  const float2 arg = __fmul2_rn(__fmul2_rn(x, x), make_float2(-0.5f * 1.4426950408889634f, -0.5f * 1.4426950408889634f));
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(arg.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(arg.y));
  const float2 erf_abs = __ffma2_rn(np, e, make_float2(1.0f, 1.0f));
  const float2 half = make_float2(0.5f, 0.5f);
  return __ffma2_rn(__fmul2_rn(ax, half), erf_abs, __fmul2_rn(x, half));
}

__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t acc) { umma_f16(d, a, b, IDESC, acc); }

// Epilogue of one 128 x 256 accumulator tile for epilogue warp (quarter, cgrp): rows m0 + 32*quarter + lane, columns
// n0 + 64*cgrp .. +64, in chunks of 16 columns (tcgen05.ld x16) to keep the register footprint of 16 epilogue warps
// under the 113-register budget of a 576-thread CTA.  Waits for the accumulator (tfull), then TMEM -> registers -> bias /
// GELU / residual / position add -> 16-byte global stores.  Shared by the 1-CTA and 2-CTA kernels.
__device__ __forceinline__ void epilogue_tile(const Params& p, const CUtensorMap* tmO, uint32_t tmem_acc, int m0, int n0,
                                              int quarter, int cgrp, int lane, float* s_bias_stage, float* s_stage,
                                              uint32_t tfull_addr, uint32_t tfull_parity) {
  const int row = m0 + quarter * 32 + lane;
  const bool row_ok = row < p.M;
  constexpr int NCH = EPI_COLS / 16;
  const int cbeg = cgrp * NCH;        // chunk index in units of 16 columns
  // EPI_RESID (fp32 residual stream, read-modify-write): one thread per ROW would touch 32 different 128-byte lines per
  // instruction, which makes the LSU — not the MMAs — the limiter of the K = 1280 out-projection.  The 32 x 16 block is
  // therefore transposed through a per-warp smem buffer so that groups of 4 lanes read / add / write the 64 contiguous
  // bytes one row owns in this chunk with 16-byte accesses (8 rows per instruction).  The residual of ALL of this
  // warp's chunks is requested before the accumulator wait: the epilogue warps would otherwise sit idle there, and with
  // one chunk in flight per warp the HBM / L2 latency (not bandwidth) set the epilogue time.
  const int rsub = lane >> 2, csub = (lane & 3) * 4;   // coalesced phase: rows 8*it + rsub, columns csub .. csub+3
  constexpr int XDEPTH = 2;  // chunks of residual in flight per warp (register budget: 168 with 10 warps per CTA)
  float4 xr[XDEPTH][4];
  float* xbase = reinterpret_cast<float*>(p.out) + (size_t)(m0 + quarter * 32) * p.ldo + csub;
  auto load_resid = [&](int cc, float4* dst) {
    const int col0 = n0 + (cbeg + cc) * 16;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int rr = 8 * it + rsub;
      dst[it] = (col0 < p.N && m0 + quarter * 32 + rr < p.M)
                    ? *reinterpret_cast<const float4*>(xbase + (size_t)rr * p.ldo + col0)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  // EPI_RESID, default: the add is not done here at all.  Each 32 x 16 block of (accumulator + bias) is written to the
  // warp's staging block and handed to the L2 as ONE fp32 TMA reduction (cp.reduce.async.bulk.tensor .add): no residual
  // loads, no load latency to hide, half the SM <-> L2 bytes, the same single rounding fl(x + fl(acc + bias)) per
  // element.  tools/micro/tma_reduce_bench.cu: bulk add.f32 over the 491 MB stream runs at the HBM roof (5.5 TB/s
  // read + write), like a fully pipelined read-modify-write.
  const bool resid_rmw = p.epi == EPI_RESID && !p.resid_tma;
  if (resid_rmw) {
    if (row_ok) {
      // the chunks that are not register-prefetched are at least pulled into L2 while the MMAs of this tile still run
      const char* rp = reinterpret_cast<const char*>(reinterpret_cast<const float*>(p.out) + (size_t)row * p.ldo + n0 +
                                                     cgrp * EPI_COLS);
#pragma unroll
      for (int j = XDEPTH * 64 / 128; j < EPI_COLS * 4 / 128; ++j)
        if (n0 + cgrp * EPI_COLS + j * 32 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + j * 128));
    }
#pragma unroll
    for (int cc = 0; cc < XDEPTH; ++cc) load_resid(cc, xr[cc]);
  }
  if (threadIdx.x < BN) {  // this tile's 256 bias values -> smem; reads below are conflict-free broadcasts
    const int col = n0 + threadIdx.x;
    s_bias_stage[threadIdx.x] = (p.bias && col < p.N) ? __ldg(p.bias + col) : 0.0f;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
  mbar_wait(tfull_addr, tfull_parity);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
  for (int cc = 0; cc < NCH; ++cc) {  // unrolled: xr[cc] must stay in registers
    const int c = cbeg + cc;
    const int col0 = n0 + c * 16;
    if (col0 >= p.N) break;  // warp-uniform
    uint32_t r[16];
    tmem_ld16(taddr + c * 16, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
    if (p.epi == EPI_RESID && p.resid_tma) {
      const float4* b4p = reinterpret_cast<const float4*>(s_bias_stage + c * 16);
      // the TMA engine has read the previous block out of this warp's staging memory
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      const int sw = (lane >> 1) & 3;  // CU_TENSOR_MAP_SWIZZLE_64B: 16-byte chunk index ^= address bits [7:8] = (row >> 1) & 3
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b4 = b4p[j];
        *reinterpret_cast<float4*>(s_stage + lane * 16 + ((j ^ sw) << 2)) =
            make_float4(v[4 * j] + b4.x, v[4 * j + 1] + b4.y, v[4 * j + 2] + b4.z, v[4 * j + 3] + b4.w);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {  // rows >= M and columns >= N are clipped by the tensor map
        tma_reduce_add_2d(tmO, smem_u32(s_stage), col0, m0 + quarter * 32);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      continue;
    }
    if (p.epi == EPI_RESID) {
      const float4* b4p = reinterpret_cast<const float4*>(s_bias_stage + c * 16);
      __syncwarp();  // previous chunk's coalesced reads of the staging buffer are done
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b4 = b4p[j];
        *reinterpret_cast<float4*>(s_stage + lane * EPI_STAGE_LD + 4 * j) =
            make_float4(v[4 * j] + b4.x, v[4 * j + 1] + b4.y, v[4 * j + 2] + b4.z, v[4 * j + 3] + b4.w);
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rr = 8 * it + rsub;
        if (m0 + quarter * 32 + rr < p.M) {
          const float4 a4 = *reinterpret_cast<const float4*>(s_stage + rr * EPI_STAGE_LD + csub);
          const float4 x4 = xr[cc % XDEPTH][it];
          *reinterpret_cast<float4*>(xbase + (size_t)rr * p.ldo + col0) =
              make_float4(x4.x + a4.x, x4.y + a4.y, x4.z + a4.z, x4.w + a4.w);
        }
      }
      if (cc + XDEPTH < NCH) load_resid(cc + XDEPTH, xr[cc % XDEPTH]);  // refill the slot just consumed
      continue;
    }
    const bool tma_st = p.out_bf16 && p.store_tma;  // warp-uniform; rows >= M are computed too and clipped by the tensor map
    if (row_ok || tma_st) {
      {
        const float4* b4p = reinterpret_cast<const float4*>(s_bias_stage + c * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b4 = b4p[j];
          v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
        }
      }
      if (p.epi == EPI_GELU || p.epi == EPI_GELU_POS) {
        if (p.out_bf16) {
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float2 g2 = gelu_fast2(make_float2(v[j], v[j + 1]));
            v[j] = g2.x;
            v[j + 1] = g2.y;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
        }
      }
      if (p.epi == EPI_GELU_POS) {
        const float* pr = p.pos + (size_t)(row % p.pos_period) * p.N + col0;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 q4 = __ldg(reinterpret_cast<const float4*>(pr + j));
          v[j] += q4.x; v[j + 1] += q4.y; v[j + 2] += q4.z; v[j + 3] += q4.w;
        }
      }
      if (tma_st) {
        // two 16-column chunks make one 32 x 32 bf16 block (64 B per row, the tensor map's 64 B swizzle) -> one TMA store
        const int half = cc & 1, sw = (lane >> 1) & 3;
        if (half == 0) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous block has left smem
          __syncwarp();
        }
        char* blk = reinterpret_cast<char*>(s_stage) + lane * 64;
        uint4 w0, w1;
        w0.x = pack_bf16(v[0], v[1]); w0.y = pack_bf16(v[2], v[3]); w0.z = pack_bf16(v[4], v[5]); w0.w = pack_bf16(v[6], v[7]);
        w1.x = pack_bf16(v[8], v[9]); w1.y = pack_bf16(v[10], v[11]); w1.z = pack_bf16(v[12], v[13]); w1.w = pack_bf16(v[14], v[15]);
        *reinterpret_cast<uint4*>(blk + (((2 * half) ^ sw) << 4)) = w0;
        *reinterpret_cast<uint4*>(blk + (((2 * half + 1) ^ sw) << 4)) = w1;
        if (half == 1) {  // N % 32 == 0 for the wide kernels: chunks always come in pairs
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(tmO, smem_u32(s_stage), col0 - 16, m0 + quarter * 32);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else if (p.out_bf16) {
        bf16* o = reinterpret_cast<bf16*>(p.out) + (size_t)row * p.ldo + col0;
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          uint4 w4;
          w4.x = pack_bf16(v[j], v[j + 1]); w4.y = pack_bf16(v[j + 2], v[j + 3]);
          w4.z = pack_bf16(v[j + 4], v[j + 5]); w4.w = pack_bf16(v[j + 6], v[j + 7]);
          *reinterpret_cast<uint4*>(o + j) = w4;
        }
      } else {
        float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col0;
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  const uint32_t bar0 = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 4);
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)) + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 4));

  float* s_bias = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + STAGES * STAGE_BYTES + 256);  // [2][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN, n_tiles = tiles_m * tiles_n;
  const int k_blocks = p.K / BK;

  if (warp == N_EPI_WARPS && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), N_EPI_WARPS * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == N_EPI_WARPS + 1) {  // one warp allocates all 512 TMEM columns (two 256-column accumulators)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == N_EPI_WARPS) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(empty_bar(s), ((it / STAGES) & 1) ^ 1);
          mbar_expect_tx(full_bar(s), STAGE_BYTES);
          const uint32_t sa = base + s * STAGE_BYTES;
          tma_load_2d(sa, &tmA, full_bar(s), kb * BK, m0);
          tma_load_2d(sa + A_BYTES, &tmB, full_bar(s), kb * BK, n0);
        }
      }
    }
  } else if (warp == N_EPI_WARPS + 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0, tcount = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tcount) {
        const uint32_t as = tcount & 1;
        mbar_wait(tempty_bar(as), ((tcount >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(full_bar(s), (it / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = base + s * STAGE_BYTES;
          const uint64_t da = make_desc(sa), db = make_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)  // +32 B per K=16 step inside the 128 B swizzle atom
            umma(d_tmem, da + (uint64_t)(k * UMMA_K * 2 >> 4), db + (uint64_t)(k * UMMA_K * 2 >> 4), (kb | k) != 0);
          umma_commit(empty_bar(s));  // smem stage reusable once these MMAs have read it
        }
        umma_commit(tfull_bar(as));   // accumulator complete
      }
    }
  } else {
    // ===================== epilogue: warp w reads TMEM lanes 32*(w%4).. and columns 64*(w/4) .. +64 =====================
    const int quarter = warp & 3, cgrp = warp >> 2;
    uint32_t tcount = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tcount) {
      const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN;
      const uint32_t as = tcount & 1;
      epilogue_tile(p, &tmO, tmem_base + as * BN, m0, n0, quarter, cgrp, lane, s_bias + as * BN,
                    s_bias + (EPI_STAGE_OFF - 256) / 4 + warp * EPI_STAGE_FLOATS, tfull_bar(as), (tcount >> 1) & 1);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tempty_bar(as));  // one arrival per epilogue thread hands the accumulator back to the MMA warp
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // this warp's TMA reductions have landed
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == N_EPI_WARPS + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a pair of CTAs on one TPC computes a 256 x 256 tile with tcgen05.mma M = 256.  Each CTA
// stages its own 128 rows of A and only HALF of the W tile (128 of the 256 weight rows), so the shared-memory fill and
// the L2 -> SM traffic per MMA drop by a third (32 KB instead of 48 KB per CTA per k-block), which is what limits the
// single-CTA kernel.  The leader CTA (cluster rank 0) issues every MMA; both CTAs run a TMA producer that credits the
// LEADER's full barrier; tcgen05.commit multicasts the "stage free" and "accumulator ready" arrivals to both CTAs; the
// non-leader's epilogue threads release the accumulator on the leader's barrier through a shared::cluster arrive.
namespace c2 {
constexpr int STAGES = 6;
constexpr int A_BYTES = 128 * BK * 2, B_BYTES = 128 * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;  // per CTA: 32 KB
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + EPI_STAGE_OFF + KW_EPI_WARPS * EPI_STAGE_FLOATS * 4;
constexpr uint32_t IDESC = make_idesc(256, BN, 0, 0);
}  // namespace c2

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmO, const Params p) {
  constexpr int STAGES = c2::STAGES, A_BYTES = c2::A_BYTES, STAGE_BYTES = c2::STAGE_BYTES;
  constexpr uint32_t IDESC = c2::IDESC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 4);
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gen_base + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 4));
  float* s_bias = reinterpret_cast<float*>(gen_base + STAGES * STAGE_BYTES + 256);  // [2][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int tiles_m = (p.M + 255) / 256, tiles_n = (p.N + BN - 1) / BN, n_tiles = tiles_m * tiles_n;
  const int k_blocks = p.K / BK;

  if (warp == N_EPI_WARPS && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);    // used in the leader only: one arrive.expect_tx for the pair's 64 KB
      mbar_init(empty_bar(s), 1);   // one multicast commit per use, in each CTA
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);                       // multicast commit, in each CTA
      mbar_init(tempty_bar(s), 2 * N_EPI_WARPS * 32);   // leader only: both CTAs' epilogue threads
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == N_EPI_WARPS + 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == N_EPI_WARPS) {
    // ===================== TMA producer (both CTAs): own 128 rows of A, own half of the W tile =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = pair; t < n_tiles; t += n_pairs) {
        const int m0 = (t / tiles_n) * 256 + (int)rank * 128, n0 = (t % tiles_n) * BN + (int)rank * 128;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(empty_bar(s), ((it / STAGES) & 1) ^ 1);
          if (rank == 0) mbar_expect_tx(full_bar(s), 2 * STAGE_BYTES);  // bytes of both CTAs land on this barrier
          const uint32_t lead_full = mapa_u32(full_bar(s), 0);
          const uint32_t sa = base + s * STAGE_BYTES;
          tma_load_2d_2sm(sa, &tmA, lead_full, kb * BK, m0);
          tma_load_2d_2sm(sa + A_BYTES, &tmB, lead_full, kb * BK, n0);
        }
      }
    }
  } else if (warp == N_EPI_WARPS + 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      uint32_t it = 0, tcount = 0;
      for (int t = pair; t < n_tiles; t += n_pairs, ++tcount) {
        const uint32_t as = tcount & 1;
        mbar_wait(tempty_bar(as), ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(full_bar(s), (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = base + s * STAGE_BYTES;
          const uint64_t da = make_desc(sa), db = make_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_f16_2sm(d_tmem, da + 2 * k, db + 2 * k, IDESC, (kb | k) != 0);
          umma_commit_2sm(empty_bar(s), 3);  // both CTAs' stage s is reusable once these MMAs have read it
        }
        umma_commit_2sm(tfull_bar(as), 3);   // accumulator complete in both CTAs
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int quarter = warp & 3, cgrp = warp >> 2;
    uint32_t tcount = 0;
    for (int t = pair; t < n_tiles; t += n_pairs, ++tcount) {
      const int m0 = (t / tiles_n) * 256 + (int)rank * 128, n0 = (t % tiles_n) * BN;
      const uint32_t as = tcount & 1;
      epilogue_tile(p, &tmO, tmem_base + as * BN, m0, n0, quarter, cgrp, lane, s_bias + as * BN,
                    s_bias + (EPI_STAGE_OFF - 256) / 4 + warp * EPI_STAGE_FLOATS, tfull_bar(as), (tcount >> 1) & 1);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive_cluster(mapa_u32(tempty_bar(as), 0));  // both CTAs' epilogue threads hand the accumulator back
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // this warp's TMA reductions have landed
  }

  tc_fence_before();
  cluster_sync_all();  // nobody frees TMEM or exits while the peer may still signal / read
  if (warp == N_EPI_WARPS + 1) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// Decode-time ("skinny") variant: M = batch <= 64 (or <= 128) rows of activations against a weight matrix that is read exactly once.
// The product is computed transposed, out^T[N, M] = W[N,K] . A[M,K]^T, so the 128-row MMA dimension streams weight
// rows and the batch is the N = 64 dimension: R x 64 x K tiles, a deep TMA ring (this kernel is bound by how fast one SM
// can pull bytes, ~50 GB/s per CTA measured), two 64-column TMEM accumulators.
//
// Split-K over a thread-block cluster: the d x d and fc2 projections (N = 1280) only make 40 tiles, so their K-loop ran
// on 40 of 148 SMs.  Launched as clusters of S CTAs, CTA `rank` of a cluster accumulates k-blocks [rank K/S, (rank+1) K/S)
// of the SAME tile and drops its raw fp32 partial into slot `rank` of the staging buffer of CTA 0 through distributed
// shared memory; after one cluster barrier CTA 0 adds the S partials in rank order (deterministic: no atomics) and runs
// the epilogue.
//
// Epilogue: the accumulator (lane = output feature, column = batch row) is transposed through the staging buffer so
// that all four epilogue warps write whole 16-byte groups of one batch row; bias / residual values are requested in that
// same layout before the accumulator wait.
namespace sk {
constexpr int BM = 128;       // MMA M: weight rows
constexpr int MAX_BN = 128;   // MMA N: batch rows, 64 or 128 (template parameter BN of the kernel)
constexpr int MAX_SPLIT = 4;
// R = weight rows actually loaded (and produced) per tile.  R = 32 quarters the bytes per stage, so small-N projections
// spread over 4x more CTAs; the MMA still runs at M = 128 and simply reads stale shared memory for rows R..127, whose
// accumulator rows are never read back.
// BN = 128 (batches of 65..128 rows: two coalesced 64-utterance batches decode as one, the per-position latency chain
// is paid once for both): activation tiles double, so the rings are shallower; everything else is the same kernel.
template <int R, int BN> struct Cfg {
  static constexpr int A_BYTES = R * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN == 64 ? (R == 32 ? 12 : R == 40 ? 10 : 8) : (R == 128 ? 4 : 8);
  // k-block slots that share one full / empty mbarrier pair (one wait + fence per group on the MMA-issuing thread)
  static constexpr int GROUP = R == 32 ? 4 : 2;  // (STAGES is a multiple of GROUP)
  static constexpr int RING = STAGES * STAGE_BYTES + (BM - R) * BK * 2;  // + tail the M = 128 read of the last stage may touch
  static constexpr int SLOTS = R == 32 ? (BN == 64 ? MAX_SPLIT : 3) : 1;  // split-K partial slots (R = 32 only)
  // [slot][batch row][feature] fp32; R = 128 (vocabulary): + one float of pitch per row, the per-row column masks and
  // 512 B of lane bits for the fused arg-max epilogue (EPI_ARGMAX)
  static constexpr int OUT_STAGE = SLOTS * BN * R * 4 + (R == 128 ? 2 * BN * 4 + 512 : 0);
  static constexpr size_t SMEM_BYTES = 1024 + (size_t)RING + 512 + OUT_STAGE;
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulators
  static constexpr uint32_t IDESC = make_idesc(BM, BN, 0, 0);
  static_assert(STAGES % GROUP == 0 && STAGE_BYTES % 1024 == 0 && A_BYTES % 1024 == 0, "ring layout");
  static_assert(SMEM_BYTES <= 232448, "skinny GEMM smem budget");
};
template struct Cfg<32, 64>; template struct Cfg<40, 64>; template struct Cfg<128, 64>;
template struct Cfg<32, 128>; template struct Cfg<40, 128>; template struct Cfg<128, 128>;
}  // namespace sk

__device__ __forceinline__ void st_cluster_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}

// Coalesced half of the skinny epilogue for one tile: V consecutive features of one batch row per item, items spread
// over the 128 epilogue threads.  `stage` = [n_slots][64][R] fp32 partial sums, `res` = residual values preloaded in the
// same item order (EPI_RESID, V = 4 only).
template <int R, int BN, int V, int EPI, bool OUT_BF16>
__device__ __forceinline__ void skinny_store(const Params& p, const float* stage, int n_slots, int n0, int tid,
                                             const float4* res, const float4& bias4) {
  constexpr int GROUPS = R / V, ITEMS = BN * GROUPS, PER = ITEMS / (SK_EPI_WARPS * 32);
  // res[i] (R = 32, V = 4: PER = BN / 16) needs the full unroll; longer loops keep registers down
  constexpr int UNROLL = (R == 32 && V == 4) ? PER : PER <= 4 ? PER : 4;
#pragma unroll UNROLL
  for (int i = 0; i < PER; ++i) {
    const int idx = tid + i * SK_EPI_WARPS * 32, row = idx / GROUPS, f = (idx % GROUPS) * V, n = n0 + f;
    if (row >= p.M || n >= p.N) continue;
    float v[V];
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = stage[row * R + f + e];
    for (int sl = 1; sl < n_slots; ++sl) {
#pragma unroll
      for (int e = 0; e < V; ++e) v[e] += stage[(sl * BN + row) * R + f + e];
    }
    if (V == 4 && R == 32) {  // every item of a thread covers the same 4 features: one bias group, requested up front
      v[0] += bias4.x; v[1] += bias4.y; v[V - 2] += bias4.z; v[V - 1] += bias4.w;
    } else if (p.bias) {
#pragma unroll
      for (int e = 0; e < V; ++e) v[e] += __ldg(p.bias + n + e);
    }
    if (EPI == EPI_GELU) {
#pragma unroll
      for (int e = 0; e < V; ++e) v[e] = OUT_BF16 ? gelu_fast(v[e]) : gelu_erf(v[e]);
    }
    if (OUT_BF16) {
      bf16* o = reinterpret_cast<bf16*>(p.out) + (size_t)row * p.ldo + n;
      if (V == 4) *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
      else if (V == 2) *reinterpret_cast<uint32_t*>(o) = pack_bf16(v[0], v[V - 1]);
      else o[0] = __float2bfloat16_rn(v[0]);
    } else {
      float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + n;
      if (EPI == EPI_RESID) {
        if (V == 4 && R == 32) { v[0] += res[i].x; v[1] += res[i].y; v[V - 2] += res[i].z; v[V - 1] += res[i].w; }
        else {
#pragma unroll
          for (int e = 0; e < V; ++e) v[e] += o[e];
        }
      }
      if (V == 4) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[V - 2], v[V - 1]);
      else if (V == 2) *reinterpret_cast<float2*>(o) = make_float2(v[0], v[V - 1]);
      else o[0] = v[0];
    }
  }
}

template <int R, int BN, int V>
__device__ __forceinline__ void skinny_store_dispatch(const Params& p, const float* stage, int n_slots, int n0, int tid,
                                                      const float4* res, const float4& bias4) {
  if (p.out_bf16) {
    if (p.epi == EPI_GELU) skinny_store<R, BN, V, EPI_GELU, true>(p, stage, n_slots, n0, tid, res, bias4);
    else skinny_store<R, BN, V, EPI_STORE, true>(p, stage, n_slots, n0, tid, res, bias4);
  } else {
    if (p.epi == EPI_GELU) skinny_store<R, BN, V, EPI_GELU, false>(p, stage, n_slots, n0, tid, res, bias4);
    else if (p.epi == EPI_RESID) skinny_store<R, BN, V, EPI_RESID, false>(p, stage, n_slots, n0, tid, res, bias4);
    else skinny_store<R, BN, V, EPI_STORE, false>(p, stage, n_slots, n0, tid, res, bias4);
  }
}

template <int R, int BN>
__global__ void __launch_bounds__(SK_THREADS, 1)
gemm_tc_skinny_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA, const Params p,
                      const int vec) {
  using C = sk::Cfg<R, BN>;
  constexpr int BM = R, STAGES = C::STAGES, A_BYTES = C::A_BYTES, STAGE_BYTES = C::STAGE_BYTES;
  constexpr int GROUP = C::GROUP;  // barrier g serves slots [g*GROUP, (g+1)*GROUP)
  constexpr int TMEM_COLS = C::TMEM_COLS;
  constexpr uint32_t IDESC = C::IDESC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + C::RING;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gen_base + C::RING + 8 * (2 * STAGES + 4));
  float* out_stage = reinterpret_cast<float*>(gen_base + C::RING + 512);  // [slot][BN][R]
  const uint32_t out_stage_u32 = base + C::RING + 512;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  auto stamp = [&](int i) {
    if (p.stamps && blockIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.stamps[i] = t;
    }
  };
  if (threadIdx.x == 0) stamp(0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // split-K: the S CTAs of a cluster share one tile
  const int S = R == 32 ? (int)cluster_nctarank() : 1, rank = R == 32 ? (int)cluster_ctarank() : 0;
  const int n_tiles = (p.N + BM - 1) / BM;  // tiles over output features
  const int n_workers = (int)gridDim.x / S, worker = (int)blockIdx.x / S;
  const int kb_all = p.K / BK;
  const int kb_begin = rank * kb_all / S, k_blocks = (rank + 1) * kb_all / S - kb_begin;  // this CTA's k-range
  const int my_tiles = (n_tiles - worker + n_workers - 1) / n_workers;

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), SK_EPI_WARPS * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // split-K handshake #1: every CTA of the cluster is running before anyone stores into CTA 0's shared memory (the
  // matching wait sits right before those stores, long after the arrival — it never stalls)
  if (R == 32 && S > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");

  if (warp == 4) {
    if (lane == 0) {
      const uint32_t total = (uint32_t)my_tiles * k_blocks;
      auto tile_of = [&](uint32_t it) { return worker + (int)(it / k_blocks) * n_workers; };
      auto kcoord = [&](uint32_t it) { return (kb_begin + (int)(it % k_blocks)) * BK; };
      // Programmatic dependent launch: weights do not depend on the preceding kernel, so the first ring of weight
      // tiles is requested before griddepcontrol.wait; only the activation tiles wait for the producer grid.
      const uint32_t pre = total < (uint32_t)STAGES ? total : (uint32_t)STAGES;
      auto group_bytes = [&](uint32_t it0) {  // bytes of the barrier group that starts at k-block it0
        const uint32_t n = total - it0 < (uint32_t)GROUP ? total - it0 : (uint32_t)GROUP;
        return n * (uint32_t)STAGE_BYTES;
      };
      const uint64_t wpol = p.w_hint == 1 ? l2_evict_last() : p.w_hint == 2 ? l2_evict_first() : 0ull;
      auto load_w = [&](uint32_t dst, uint32_t bar, int c0, int c1) {
        if (p.w_hint) tma_load_2d_hint(dst, &tmW, bar, c0, c1, wpol);
        else tma_load_2d(dst, &tmW, bar, c0, c1);
      };
      for (uint32_t it = 0; it < pre; ++it) {
        const int g = (int)it / GROUP;
        if (it % GROUP == 0) mbar_expect_tx(full_bar(g), group_bytes(it));
        load_w(base + it * STAGE_BYTES, full_bar(g), kcoord(it), tile_of(it) * BM);
      }
      stamp(1);
      asm volatile("griddepcontrol.wait;" ::: "memory");
      stamp(2);
      for (uint32_t it = 0; it < pre; ++it)
        tma_load_2d(base + it * STAGE_BYTES + A_BYTES, &tmA, full_bar((int)it / GROUP), kcoord(it), 0);
      for (uint32_t it = pre; it < total; ++it) {
        const int s = it % STAGES, g = s / GROUP;
        if (it % GROUP == 0) {
          mbar_wait(empty_bar(g), ((it / STAGES) & 1) ^ 1);
          mbar_expect_tx(full_bar(g), group_bytes(it));
        }
        const uint32_t sa = base + s * STAGE_BYTES;
        load_w(sa, full_bar(g), kcoord(it), tile_of(it) * BM);
        tma_load_2d(sa + A_BYTES, &tmA, full_bar(g), kcoord(it), 0);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      uint32_t it = 0;
      const uint32_t total_it = (uint32_t)my_tiles * (uint32_t)k_blocks;
      for (uint32_t tcount = 0; tcount < (uint32_t)my_tiles; ++tcount) {
        const uint32_t as = tcount & 1;
        mbar_wait(tempty_bar(as), ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % STAGES, g = s / GROUP;
          if (it % GROUP == 0) {
            mbar_wait(full_bar(g), (it / STAGES) & 1);
            if (it == 0) stamp(3);
            tc_fence_after();
          }
          const uint32_t sa = base + s * STAGE_BYTES;
          const uint64_t da = make_desc(sa), db = make_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_f16(d_tmem, da + 2 * k, db + 2 * k, IDESC, (kb | k) != 0);
          if (it % GROUP == GROUP - 1 || it + 1 == total_it) umma_commit(empty_bar(g));  // group consumed
        }
        umma_commit(tfull_bar(as));
        stamp(4);
        if (tcount == 0) stamp(11);
      }
    }
  } else {
    asm volatile("griddepcontrol.wait;" ::: "memory");  // outputs / residual belong to the preceding kernels
    const int tid = threadIdx.x;                  // 0..127
    const bool row_warp = warp * 32 < R;          // TMEM lanes R..127 hold garbage rows
    // slot `rank` of the staging buffer of the cluster's CTA 0 (this CTA's own buffer when there is no split)
    const uint32_t my_slot = out_stage_u32 + (uint32_t)rank * (BN * R * 4);
    const uint32_t slot_remote = S > 1 ? mapa_u32(my_slot, 0) : 0u;
    // EPI_ARGMAX: what the rules need to know about a batch row to mask TEXT ids (sr::text_col_mask), once per CTA from
    // the last two tokens of the history.  Layout of the staging area: [64 column masks][128 lane bits][64 x 129 floats].
    int* s_cm = reinterpret_cast<int*>(out_stage);
    int* s_lb = s_cm + BN;
    float* s_tile = reinterpret_cast<float*>(s_lb + 128);
    constexpr int TP = 129;  // pitch of a batch row in s_tile: conflict-free both for the transposing writes and the scans
    if (R == 128 && p.epi == EPI_ARGMAX) {
      if (tid < BN) {
        int cm = 0xf;  // rows past the batch: everything masked
        if (tid < p.M) {
          const int* trow = p.sf.tokens + (size_t)tid * p.sf.ld_tokens;
          const int n = p.sf.pos + 1 - p.sf.begin_index;
          const int t0 = n >= 1 ? trow[p.sf.pos] : 0, t1 = n >= 2 ? trow[p.sf.pos - 1] : 0;
          cm = sr::text_col_mask(p.sf.return_ts, n, t0, t1, p.sf.rules.ts_begin);
        }
        s_cm[tid] = cm;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    for (int tcount = 0; tcount < my_tiles; ++tcount) {
      const int t = worker + tcount * n_workers;
      const uint32_t as = tcount & 1;
      if (R == 128 && p.epi == EPI_ARGMAX) {
        // Fused logits processors + arg-max.  The accumulator (lane = vocabulary row, column = batch row) is transposed
        // through shared memory; then thread (batch row, half tile) scans its 64 logits sequentially — masking is one AND
        // of the id's lane bits with the row's column bits, ties keep the smaller id as torch.argmax does, and no
        // warp-collective sits on the path (a cross-lane reduction per batch column cost ~8 us per tile).  Each thread
        // emits ONE (best value, id) pair; the few 32-row slices that reach into the timestamp ids additionally leave
        // the kernel as raw fp32 logits (1.5 k per row) — their rules (pairing, monotonicity, logsumexp-vs-max) need the
        // whole timestamp range and run in sample_combine_kernel.  The 13 MB fp32 logit matrix is never written.
        const int lo = t * BM + warp * 32, v = lo + lane;
        const bool valid = v < p.N, tail = lo >= p.sf.tail0;
        // this lane's id against the row-independent half of the rules (sr::text_lane_bits); 0xf = masked for every row
        s_lb[warp * 32 + lane] = (valid && !tail) ? (int)sr::text_lane_bits(p.sf.rules, p.sf.return_ts, v, p.sf.flags[v]) : 0xf;
        mbar_wait(tfull_bar(as), (tcount >> 1) & 1);
        if (threadIdx.x == 0 && tcount == 0) stamp(8);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + as * BN;
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          if (c * 32 >= p.M) break;
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (tail && valid) {
            float* o = p.sf.tail + (size_t)(c * 32) * p.sf.tail_ld + (v - p.sf.tail0);
            const int nb = min(32, p.M - c * 32);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nb) o[(size_t)j * p.sf.tail_ld] = __uint_as_float(r[j]);
          } else if (!tail) {
#pragma unroll
            for (int j = 0; j < 32; ++j) s_tile[(c * 32 + j) * TP + warp * 32 + lane] = __uint_as_float(r[j]);
          }
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(as));
        asm volatile("bar.sync 1, 128;" ::: "memory");  // tile + lane bits staged
        if (t * BM < p.sf.tail0) {                      // tiles that hold text ids emit partials
          // BN = 64: thread = (batch row, half tile); BN = 128: thread = batch row, both halves in turn
#pragma unroll
          for (int hh = 0; hh < BN / 64; ++hh) {
            const int b = BN == 64 ? (tid & 63) : tid, half = BN == 64 ? (tid >> 6) : hh, cm = s_cm[b];
            const float* row = s_tile + b * TP + half * 64;
            const int* lbs = s_lb + half * 64;
            float best = -INFINITY;
            int bi = -1;
#pragma unroll 16
            for (int k = 0; k < 64; ++k) {
              const float x = row[k];
              const bool take = !(lbs[k] & cm) && x > best;  // strict: the first (smallest) id wins a tie
              best = take ? x : best;
              bi = take ? k : bi;
            }
            if (b < p.M)
              p.sf.vpart[(size_t)b * p.sf.n_part + t * 2 + half] =
                  make_float2(best, __int_as_float(bi >= 0 ? t * BM + half * 64 + bi : p.N));
          }
        }
        if (tcount + 1 < my_tiles) asm volatile("bar.sync 1, 128;" ::: "memory");  // staging free for the next tile
        if (threadIdx.x == 0) stamp(tcount == 0 ? 9 : (tcount + 1 < my_tiles ? 5 : 6));
        continue;
      }
      // EPI_RESID: this thread's residual groups are requested before the accumulator wait (R = 32: BN / 16 x 16 bytes)
      float4 res[R == 32 ? BN / 16 : 1];
      if (R == 32 && p.epi == EPI_RESID && vec == 4 && rank == 0) {
#pragma unroll
        for (int i = 0; i < BN / 16; ++i) {
          const int idx = tid + i * 128, row = idx / 8, n = t * BM + (idx % 8) * 4;
          res[i] = (row < p.M && n < p.N)
                       ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.out) + (size_t)row * p.ldo + n)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);  // R = 32, vec = 4: this thread's only bias group (idx % 8 = tid % 8)
      if (R == 32 && vec == 4 && rank == 0 && p.bias && t * BM + (tid % 8) * 4 < p.N)
        bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + t * BM + (tid % 8) * 4));
      mbar_wait(tfull_bar(as), (tcount >> 1) & 1);
      if (threadIdx.x == 0 && tcount == 0) stamp(8);
      tc_fence_after();
      if (S > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");  // handshake #1
      if (row_warp) {
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + as * BN;
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          if (c * 32 >= p.M) break;  // warp-uniform: no batch rows in this half
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int col = warp * 32 + lane;  // feature inside the tile
          if (col >= R) continue;            // R = 40: lanes 8..31 of warp 1 sit on garbage accumulator rows
          if (S > 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) st_cluster_f32(slot_remote + ((c * 32 + j) * R + col) * 4, __uint_as_float(r[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) out_stage[(c * 32 + j) * R + col] = __uint_as_float(r[j]);
          }
        }
      }
      if (threadIdx.x == 0 && tcount == 0) stamp(9);
      tc_fence_before();
      mbar_arrive(tempty_bar(as));
      if (S == 1) {
        asm volatile("bar.sync 1, 128;" ::: "memory");  // staging buffer complete
        if (vec == 4) skinny_store_dispatch<R, BN, 4>(p, out_stage, 1, t * BM, tid, res, bias4);
        else if (vec == 2) skinny_store_dispatch<R, BN, 2>(p, out_stage, 1, t * BM, tid, res, bias4);
        else skinny_store_dispatch<R, BN, 1>(p, out_stage, 1, t * BM, tid, res, bias4);
        if (tcount + 1 < my_tiles) asm volatile("bar.sync 1, 128;" ::: "memory");  // buffer free for the next tile
        if (threadIdx.x == 0) stamp(5 + (tcount == 0 ? 0 : 1));
      } else {
        // split-K (one tile per cluster): every CTA's partial lands in CTA 0, which finishes the tile
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        if (rank == 0) {
          asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
          if (threadIdx.x == 0) stamp(10);
          if (vec == 4) skinny_store_dispatch<R, BN, 4>(p, out_stage, S, t * BM, tid, res, bias4);
          else if (vec == 2) skinny_store_dispatch<R, BN, 2>(p, out_stage, S, t * BM, tid, res, bias4);
          else skinny_store_dispatch<R, BN, 1>(p, out_stage, S, t * BM, tid, res, bias4);
          if (threadIdx.x == 0) stamp(5);
        }
      }
    }
  }
  if (R == 32 && S > 1 && warp >= 4) {
    // the TMA / MMA warps take part in the cluster barrier too (it counts every thread of every CTA)
    __syncwarp();
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");  // handshake #1
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  }
  if (R == 32 && S > 1 && !(rank == 0 && warp < 4)) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(7);
  if (warp == 5) tmem_dealloc(tmem_base, TMEM_COLS);
}

static unsigned long long* g_stamps = nullptr;
static int g_use_2cta = 1;

// 2D bf16 row-major [rows, cols] (ld elements between rows) -> tiles of box_rows x 64 columns
static int make_map(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  return make_map_bf16(map, ptr, 2, gdim, gstride, box);
}

}  // namespace tc

void gemm_tc_set_stamps(unsigned long long* p) { tc::g_stamps = p; }
void gemm_tc_set_2cta(int on) { tc::g_use_2cta = on; }

int gemm_tc(const GemmArgs& g, cudaStream_t st) {
  using namespace tc;
  if (g.a_type != KW_BF16 || g.w_type != KW_BF16) return KW_ERR_UNSUPPORTED;
  const bool skinny = g.M <= sk::MAX_BN && g.epi != EPI_GELU_POS;
  if (g.K % BK != 0 || g.lda % 8 != 0 || g.M < 1) return KW_ERR_UNSUPPORTED;
  if (!skinny && (g.N % 32 != 0 || g.ldo % 8 != 0 || ((uintptr_t)g.out & 15))) return KW_ERR_UNSUPPORTED;
  if (((uintptr_t)g.A & 15) || ((uintptr_t)g.W & 15)) return KW_ERR_UNSUPPORTED;
  if ((g.epi == EPI_RESID || g.epi == EPI_GELU_POS) && g.out_type != KW_F32) return KW_ERR_UNSUPPORTED;
  static int n_sm = 0;
  static bool attr = false;
  if (!attr) {
    int dev = 0;
    KW_CUDA_OK(cudaGetDevice(&dev));
    KW_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    KW_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
#define KW_SK_ATTR(R_, BN_)                                                                                    \
  KW_CUDA_OK(cudaFuncSetAttribute(gemm_tc_skinny_kernel<R_, BN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)sk::Cfg<R_, BN_>::SMEM_BYTES))
    KW_SK_ATTR(32, 64); KW_SK_ATTR(40, 64); KW_SK_ATTR(128, 64);
    KW_SK_ATTR(32, 128); KW_SK_ATTR(40, 128); KW_SK_ATTR(128, 128);
#undef KW_SK_ATTR
    attr = true;
  }
  Params p;
  p.stamps = g_stamps;
  p.bias = g.bias; p.out = g.out; p.pos = g.pos;
  p.M = g.M; p.N = g.N; p.K = g.K; p.ldo = g.ldo; p.pos_period = g.pos_period > 0 ? g.pos_period : 1;
  p.epi = g.epi; p.out_bf16 = g.out_type == KW_BF16;
  memset(&p.sf, 0, sizeof(p.sf));
  p.w_hint = g.w_hint;
  static const int resid_tma = [] {  // KW_RESID_TMA=0: read-modify-write residual epilogue (A/B measurements)
    const char* e = getenv("KW_RESID_TMA");
    return e ? atoi(e) : 1;
  }();
  p.resid_tma = 0;
  p.store_tma = 0;
  static const int store_tma = [] {  // KW_STORE_TMA=0: per-thread stores of the bf16 epilogues (A/B measurements)
    const char* e = getenv("KW_STORE_TMA");
    return e ? atoi(e) : 1;
  }();
  if (g.epi == EPI_ARGMAX) {
    if (!g.sample || !(g.M <= sk::MAX_BN) || g.N <= 8192) return KW_ERR_UNSUPPORTED;  // R = 128 decode-time kernel only
    if (g.sample->tail0 % 32 != 0 || g.sample->n_part != 2 * ceil_div(g.sample->tail0, 128) || g.sample->tail0 > g.N ||
        g.sample->tail_ld < g.N - g.sample->tail0)
      return KW_ERR_ARG;
    p.sf = *g.sample;
  }
  CUtensorMap tmA, tmB, tmO;
  // wide kernels, fp32 residual epilogue: the output doubles as the target of TMA reductions (16-column blocks)
  auto make_out_map = [&]() -> int {
    p.resid_tma = g.epi == EPI_RESID && resid_tma && g.N % 4 == 0 && g.ldo % 4 == 0 && ((uintptr_t)g.out & 15) == 0;
    if (p.resid_tma) return make_map_f32_sw64(&tmO, g.out, g.M, g.N, g.ldo);
    p.store_tma = store_tma && g.out_type == KW_BF16 && (g.epi == EPI_STORE || g.epi == EPI_GELU) && g.N % 32 == 0 &&
                  g.ldo % 8 == 0 && ((uintptr_t)g.out & 15) == 0;
    if (p.store_tma) return make_map_bf16_sw64(&tmO, g.out, g.M, g.N, g.ldo);
    tmO = tmA;  // never dereferenced
    return KW_OK;
  };
  if (skinny) {  // decode-time shape: weights stream through the 128-row dimension, the batch is the MMA's N (64 or 128)
    const int bn = g.M <= 64 ? 64 : 128;
    // small projections: 32 weight rows per CTA tile -> 4x the CTAs in flight; 40 rows when 32-row tiles would spill
    // into a second wave by a few tiles (fc1: N = 5120 -> 160 tiles on 148 SMs took 15.3 us, 128 tiles of 40 rows one wave)
    int R = g.N <= 8192 ? 32 : 128;
    if (R == 32 && ceil_div(g.N, 32) > n_sm && ceil_div(g.N, 40) <= n_sm) R = 40;
    int rc = make_map(&tmB, g.W, g.N, g.K, g.K, R);
    if (rc) return rc;
    if ((rc = make_map(&tmA, g.A, g.M, g.K, g.lda, bn))) return rc;
    const int tiles = ceil_div(g.N, R);
    static const int pdl_gemm = [] {
      const char* e = getenv("KW_PDL_GEMM");
      return e ? atoi(e) : 1;
    }();
    static const int max_split = [] {  // KW_SPLITK=1 turns the cluster split-K off (A/B measurements)
      const char* e = getenv("KW_SPLITK");
      return e ? std::max(1, std::min(atoi(e), sk::MAX_SPLIT)) : 3;
    }();
    // split-K over a cluster when the tiles alone leave most SMs idle (N = 1280 projections: 40 tiles -> 120 CTAs)
    int split = 1;
    if (R == 32)
      for (int s2 = std::min(max_split, bn == 64 ? sk::Cfg<32, 64>::SLOTS : sk::Cfg<32, 128>::SLOTS); s2 > 1; --s2)
        if (tiles * s2 <= n_sm && g.K / BK >= 4 * s2) { split = s2; break; }
    // widest store the output layout allows: V features of one batch row per 16 / 8 / 4-byte (fp32) store
    const int esz = g.out_type == KW_BF16 ? 2 : 4;
    int vec = 1;
    if (g.N % 4 == 0 && g.ldo % 4 == 0 && ((uintptr_t)g.out % (4 * esz)) == 0 && ((uintptr_t)g.bias % 16) == 0) vec = 4;
    else if (g.N % 2 == 0 && g.ldo % 2 == 0 && ((uintptr_t)g.out % (2 * esz)) == 0) vec = 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(split > 1 ? tiles * split : std::min(tiles, n_sm));
    cfg.blockDim = dim3(SK_THREADS);
    cfg.dynamicSmemBytes = bn == 64 ? (R == 32 ? sk::Cfg<32, 64>::SMEM_BYTES : R == 40 ? sk::Cfg<40, 64>::SMEM_BYTES : sk::Cfg<128, 64>::SMEM_BYTES)
                                    : (R == 32 ? sk::Cfg<32, 128>::SMEM_BYTES : R == 40 ? sk::Cfg<40, 128>::SMEM_BYTES : sk::Cfg<128, 128>::SMEM_BYTES);
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    int na = 0;
    if (pdl_gemm) {  // prologue + weight prefetch overlap the preceding kernel's tail (PDL)
      attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attrs[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    if (split > 1) {
      attrs[na].id = cudaLaunchAttributeClusterDimension;
      attrs[na].val.clusterDim.x = split;
      attrs[na].val.clusterDim.y = 1;
      attrs[na].val.clusterDim.z = 1;
      ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    if (bn == 64) {
      if (R == 32) KW_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_skinny_kernel<32, 64>, tmB, tmA, p, vec));
      else if (R == 40) KW_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_skinny_kernel<40, 64>, tmB, tmA, p, vec));
      else KW_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_skinny_kernel<128, 64>, tmB, tmA, p, vec));
    } else {
      if (R == 32) KW_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_skinny_kernel<32, 128>, tmB, tmA, p, vec));
      else if (R == 40) KW_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_skinny_kernel<40, 128>, tmB, tmA, p, vec));
      else KW_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_skinny_kernel<128, 128>, tmB, tmA, p, vec));
    }
    KW_LAUNCH_OK();
    ++g_launches;
    return KW_OK;
  }
  if (g_use_2cta && g.M >= 256 && g.K >= 1024) {  // short-K tiles are epilogue-bound: 1-CTA kernel
    int rc = make_map(&tmA, g.A, g.M, g.K, g.lda, 128);
    if (rc) return rc;
    if ((rc = make_map(&tmB, g.W, g.N, g.K, g.K, 128))) return rc;
    static bool attr2 = false;
    if (!attr2) {
      KW_CUDA_OK(cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c2::SMEM_BYTES));
      attr2 = true;
    }
    const int tiles = ceil_div(g.M, 256) * ceil_div(g.N, BN);
    const int pairs = std::min(tiles, n_sm / 2);
    if ((rc = make_out_map())) return rc;
    gemm_tc2_kernel<<<2 * pairs, THREADS, c2::SMEM_BYTES, st>>>(tmA, tmB, tmO, p);
    KW_LAUNCH_OK();
    ++g_launches;
    return KW_OK;
  }
  int rc = make_map(&tmA, g.A, g.M, g.K, g.lda, BM);
  if (rc) return rc;
  rc = make_map(&tmB, g.W, g.N, g.K, g.K, BN);
  if (rc) return rc;
  const int n_tiles = ceil_div(g.M, BM) * ceil_div(g.N, BN);
  const int grid = std::min(n_tiles, n_sm);
  if ((rc = make_out_map())) return rc;
  gemm_tc_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(tmA, tmB, tmO, p);
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

}  // namespace kw
