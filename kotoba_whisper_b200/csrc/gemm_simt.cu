// CUDA-core GEMM with fused epilogues:  out[M,N] = epi(A[M,K] . W[N,K]^T + bias).
//
// This is the exact-fp32 path (fp32 "bit-identical tokens" mode: fp32 operands, fp32 FMA accumulation, no TF32) and the
// shape-agnostic fallback for operand/shape combinations the tcgen05 kernel (gemm_tc.cu) does not take.  Operands may
// be stored as fp32 or bf16 (template), arithmetic is always fp32.
//
// Tiling: BM x BN x 16 CTA tile, 256 threads, (BM/16) x (BN/16) outputs per thread with rows/cols interleaved by 16 so
// that shared-memory reads are 16-lane contiguous, A/W tiles stored k-major ([k][m]) and double buffered through
// registers.  BM = 128 for the encoder (M = 1500 B), BM = 32 for decode-time calls (M = B <= 64).
#include <atomic>

#include "common.cuh"

namespace kw {

extern std::atomic<long long> g_launches;

constexpr int GS_BK = 16;

template <typename OT>
__device__ __forceinline__ void epilogue_store(const GemmArgs& g, int row, int col, float acc) {
  float v = acc + (g.bias ? g.bias[col] : 0.0f);
  OT* o = reinterpret_cast<OT*>(g.out) + (size_t)row * g.ldo + col;
  switch (g.epi) {
    case EPI_STORE: st_f(o, v); break;
    case EPI_GELU: st_f(o, gelu_erf(v)); break;
    case EPI_RESID: st_f(o, ld_f(o) + v); break;
    case EPI_GELU_POS: st_f(o, gelu_erf(v) + g.pos[(size_t)(row % g.pos_period) * g.N + col]); break;
  }
}

template <typename AT, typename WT, typename OT, int BM, int BN>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmArgs g) {
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int A_VEC = BM * GS_BK / 4, W_VEC = BN * GS_BK / 4;  // 4-element vectors per tile
  constexpr int A_PER = (A_VEC + 255) / 256, W_PER = (W_VEC + 255) / 256;
  __shared__ __align__(16) float As[2][GS_BK][BM + 4];
  __shared__ __align__(16) float Ws[2][GS_BK][BN + 4];

  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const AT* A = reinterpret_cast<const AT*>(g.A);
  const WT* W = reinterpret_cast<const WT*>(g.W);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  float4 ra[A_PER], rw[W_PER];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int p = 0; p < A_PER; ++p) {
      int v = tid + p * 256;
      ra[p] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < A_VEC) {
        int r = v / (GS_BK / 4), kc = (v % (GS_BK / 4)) * 4;
        if (m0 + r < g.M) ra[p] = ld4(A + (size_t)(m0 + r) * g.lda + k0 + kc);
      }
    }
#pragma unroll
    for (int p = 0; p < W_PER; ++p) {
      int v = tid + p * 256;
      rw[p] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < W_VEC) {
        int r = v / (GS_BK / 4), kc = (v % (GS_BK / 4)) * 4;
        if (n0 + r < g.N) rw[p] = ld4(W + (size_t)(n0 + r) * g.K + k0 + kc);
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int p = 0; p < A_PER; ++p) {
      int v = tid + p * 256;
      if (v < A_VEC) {
        int r = v / (GS_BK / 4), kc = (v % (GS_BK / 4)) * 4;
        As[buf][kc + 0][r] = ra[p].x; As[buf][kc + 1][r] = ra[p].y;
        As[buf][kc + 2][r] = ra[p].z; As[buf][kc + 3][r] = ra[p].w;
      }
    }
#pragma unroll
    for (int p = 0; p < W_PER; ++p) {
      int v = tid + p * 256;
      if (v < W_VEC) {
        int r = v / (GS_BK / 4), kc = (v % (GS_BK / 4)) * 4;
        Ws[buf][kc + 0][r] = rw[p].x; Ws[buf][kc + 1][r] = rw[p].y;
        Ws[buf][kc + 2][r] = rw[p].z; Ws[buf][kc + 3][r] = rw[p].w;
      }
    }
  };

  const int nk = g.K / GS_BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * GS_BK);
#pragma unroll
    for (int k = 0; k < GS_BK; ++k) {
      float a[TM], w[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[buf][k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < TN; ++j) w[j] = Ws[buf][k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = m0 + ty + 16 * i;
    if (row >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = n0 + tx + 16 * j;
      if (col < g.N) epilogue_store<OT>(g, row, col, acc[i][j]);
    }
  }
}

template <typename AT, typename WT, typename OT>
static int launch_simt(const GemmArgs& g, cudaStream_t st) {
  if (g.M > 64) {
    dim3 grid(ceil_div(g.N, 128), ceil_div(g.M, 128));
    gemm_simt_kernel<AT, WT, OT, 128, 128><<<grid, 256, 0, st>>>(g);
  } else {
    dim3 grid(ceil_div(g.N, 64), ceil_div(g.M, 32));
    gemm_simt_kernel<AT, WT, OT, 32, 64><<<grid, 256, 0, st>>>(g);
  }
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

int gemm_simt(const GemmArgs& g, cudaStream_t st) {
  KW_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0 && g.K % GS_BK == 0, "gemm_simt: bad shape M=%d N=%d K=%d (K %% 16)", g.M,
             g.N, g.K);
  KW_REQUIRE(g.lda % 4 == 0, "gemm_simt: lda=%d must be a multiple of 4", g.lda);
  KW_REQUIRE(!(g.epi == EPI_RESID || g.epi == EPI_GELU_POS) || g.out_type == KW_F32,
             "gemm_simt: residual epilogues write the fp32 stream");
  const int key = (g.a_type << 2) | (g.w_type << 1) | g.out_type;
  switch (key) {
    case (KW_F32 << 2) | (KW_F32 << 1) | KW_F32: return launch_simt<float, float, float>(g, st);
    case (KW_F32 << 2) | (KW_BF16 << 1) | KW_F32: return launch_simt<float, bf16, float>(g, st);
    case (KW_BF16 << 2) | (KW_BF16 << 1) | KW_BF16: return launch_simt<bf16, bf16, bf16>(g, st);
    case (KW_BF16 << 2) | (KW_BF16 << 1) | KW_F32: return launch_simt<bf16, bf16, float>(g, st);
    case (KW_F32 << 2) | (KW_BF16 << 1) | KW_BF16: return launch_simt<float, bf16, bf16>(g, st);
    default:
      set_error("gemm_simt: unsupported dtype combination a=%d w=%d out=%d", g.a_type, g.w_type, g.out_type);
      return KW_ERR_UNSUPPORTED;
  }
}

}  // namespace kw
