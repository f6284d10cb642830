// Small memory-bound kernels around the GEMMs: conv-stem im2col, LayerNorm, decoder embedding.
// All take fp32 or bf16 storage (template T) and compute in fp32.
#include <algorithm>
#include <atomic>

#include "common.cuh"
#include "ln_row.cuh"

namespace kw {

extern std::atomic<long long> g_launches;

// ---- conv1 im2col ------------------------------------------------------------------------------------------------
// mel [B, C, T] f32 (time contiguous)  ->  A1 [B*T, 3*C],  A1[(b,t), tap*C + c] = mel[b, c, t - 1 + tap]  (0 outside)
// which makes conv1d(k=3, pad=1) a GEMM with conv1.weight repacked to [d, tap*C + c]  (modeling_whisper.py:619).
constexpr int IC_TT = 32;  // frames per CTA

template <typename T>
__global__ void __launch_bounds__(256) im2col_conv1_kernel(const float* __restrict__ mel, T* __restrict__ A1, int C,
                                                           int Tn) {
  extern __shared__ float s_mel[];  // [C][IC_TT + 3] (+1 pad keeps the column reads conflict-free)
  const int b = blockIdx.y, t0 = blockIdx.x * IC_TT;
  const int W = IC_TT + 3;
  const float* src = mel + (size_t)b * C * Tn;
  for (int i = threadIdx.x; i < C * (IC_TT + 2); i += blockDim.x) {
    int c = i / (IC_TT + 2), j = i % (IC_TT + 2);
    int t = t0 - 1 + j;
    s_mel[c * W + j] = (t >= 0 && t < Tn) ? src[(size_t)c * Tn + t] : 0.0f;
  }
  __syncthreads();
  const int K = 3 * C;
  for (int i = threadIdx.x; i < IC_TT * K; i += blockDim.x) {
    int tl = i / K, k = i % K;
    int tap = k / C, c = k % C;
    int t = t0 + tl;
    if (t < Tn) st_f(A1 + ((size_t)b * Tn + t) * K + k, s_mel[c * W + tl + tap]);
  }
}

// ---- conv2 im2col ------------------------------------------------------------------------------------------------
// h0 [B*Tin, d] (time-major rows)  ->  A2 [B*Tout, 3*d],  A2[(b,t), tap*d + c] = h0[b, 2t - 1 + tap, c]
// (conv1d k=3, stride 2, pad 1; modeling_whisper.py:620).  Pure 16-byte row copies.
template <typename T>
__global__ void __launch_bounds__(256) im2col_conv2_kernel(const T* __restrict__ h0, T* __restrict__ A2, int d, int Tin,
                                                           int Tout, size_t total_vec) {
  constexpr int VE = 16 / sizeof(T);  // elements per 16 B
  const int dv = d / VE;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < total_vec; i += stride) {
    int cv = (int)(i % dv);
    size_t r = i / dv;
    int tap = (int)(r % 3);
    size_t bt = r / 3;
    int t = (int)(bt % Tout);
    size_t b = bt / Tout;
    int tin = 2 * t - 1 + tap;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (tin >= 0 && tin < Tin) v = reinterpret_cast<const uint4*>(h0 + (b * Tin + tin) * d)[cv];
    reinterpret_cast<uint4*>(A2 + bt * 3 * (size_t)d + (size_t)tap * d)[cv] = v;
  }
}

// ---- LayerNorm ---------------------------------------------------------------------------------------------------
// rows of fp32 x -> (x - mean) * rsqrt(var + 1e-5) * w + b, biased variance, two-pass in registers; one warp per row.
// NV = float4 groups per lane.  d == NV * 128 instantiations (EXACT) carry no bounds checks and exactly NV * 4 data
// registers, which keeps the kernel at <= 64 registers -> 32 resident warps per SM (the 737 MB encoder pass is a pure
// HBM stream and needs the loads of many rows in flight); the NV = 16 generic instantiation covers any d <= 2048.
constexpr int LN_MAX_VEC = 16;

template <typename T, int NV, bool EXACT>
__global__ void __launch_bounds__(256, EXACT && NV <= 10 ? 4 : 2)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 T* __restrict__ out, int rows, int d) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (size_t)row * d;
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (EXACT || c < d) v[i] = *reinterpret_cast<const float4*>(xr + c);
  }
  ln_row<T, NV, EXACT>(v, w, bias, out + (size_t)row * d, d, lane);
}

// ---- decoder embedding: x[b] = E[tokens[b, pos]] + P[pos]  (modeling_whisper.py:738, 755-763) -----------------------
template <typename T>
__global__ void __launch_bounds__(128) embed_kernel(const int* __restrict__ tokens, int ld_tokens, int pos,
                                                    const T* __restrict__ E, const float* __restrict__ P,
                                                    float* __restrict__ x, int d, int vocab) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  int tok = tokens[(size_t)b * ld_tokens + pos];
  tok = min(max(tok, 0), vocab - 1);
  const T* e = E + (size_t)tok * d;
  const float* p = P + (size_t)pos * d;
  for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
    float4 a = ld4(e + c), q = *reinterpret_cast<const float4*>(p + c);
    *reinterpret_cast<float4*>(x + (size_t)b * d + c) = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
  }
}

// teacher-forcing variant: x[(b, t)] = E[tokens[b, t]] + P[t] for t < T
template <typename W>
__global__ void __launch_bounds__(128) embed_seq_kernel(const int* __restrict__ tokens, int ld_tokens, int T,
                                                        const W* __restrict__ E, const float* __restrict__ P,
                                                        float* __restrict__ x, int d, int vocab) {
  const int b = blockIdx.x / T, t = blockIdx.x % T;
  int tok = tokens[(size_t)b * ld_tokens + t];
  tok = min(max(tok, 0), vocab - 1);
  const W* e = E + (size_t)tok * d;
  const float* p = P + (size_t)t * d;
  float* xr = x + (size_t)blockIdx.x * d;
  for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
    float4 a = ld4(e + c), q = *reinterpret_cast<const float4*>(p + c);
    *reinterpret_cast<float4*>(xr + c) = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
  }
}

template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ in, TO* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) st_f(out + i, ld_f(in + i));
}

// ---- launchers ---------------------------------------------------------------------------------------------------
int im2col_conv1(const float* mel, void* A1, int B, int C, int Tn, kw_dtype t, cudaStream_t st) {
  dim3 grid(ceil_div(Tn, IC_TT), B);
  size_t smem = sizeof(float) * C * (IC_TT + 3);
  if (t == KW_BF16) im2col_conv1_kernel<bf16><<<grid, 256, smem, st>>>(mel, (bf16*)A1, C, Tn);
  else im2col_conv1_kernel<float><<<grid, 256, smem, st>>>(mel, (float*)A1, C, Tn);
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

int im2col_conv2(const void* h0, void* A2, int B, int d, int Tin, int Tout, kw_dtype t, cudaStream_t st) {
  const int ve = t == KW_BF16 ? 8 : 4;
  KW_REQUIRE(d % ve == 0, "im2col_conv2: d=%d not a multiple of %d", d, ve);
  size_t total = (size_t)B * Tout * 3 * (d / ve);
  int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 32);
  if (t == KW_BF16) im2col_conv2_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)h0, (bf16*)A2, d, Tin, Tout, total);
  else im2col_conv2_kernel<float><<<blocks, 256, 0, st>>>((const float*)h0, (float*)A2, d, Tin, Tout, total);
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

int layernorm(const float* x, const float* w, const float* b, void* out, int rows, int d, kw_dtype t, cudaStream_t st) {
  KW_REQUIRE(d % 4 == 0 && d <= LN_MAX_VEC * 128, "layernorm: d=%d unsupported", d);
  int blocks = ceil_div(rows, 8);
#define KW_LN_LAUNCH(TT, NV, EX)                                                                                      \
  KW_CUDA_OK(launch_pdl(PDL_LN, layernorm_kernel<TT, NV, EX>, dim3(blocks), dim3(256), 0, st, x, w, b, (TT*)out, rows, d))
  if (t == KW_BF16) {
    if (d == 1280) KW_LN_LAUNCH(bf16, 10, true);       // large / kotoba / teacher
    else if (d == 384) KW_LN_LAUNCH(bf16, 3, true);    // tiny
    else KW_LN_LAUNCH(bf16, LN_MAX_VEC, false);
  } else {
    if (d == 1280) KW_LN_LAUNCH(float, 10, true);
    else if (d == 384) KW_LN_LAUNCH(float, 3, true);
    else KW_LN_LAUNCH(float, LN_MAX_VEC, false);
  }
#undef KW_LN_LAUNCH
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

int embed(const int* tokens, int ld_tokens, int pos, const void* E, const float* P, float* x, int B, int d, int vocab,
          kw_dtype t, cudaStream_t st) {
  if (t == KW_BF16)
    KW_CUDA_OK(launch_pdl(PDL_EMBED, embed_kernel<bf16>, dim3(B), dim3(128), 0, st, tokens, ld_tokens, pos, (const bf16*)E, P, x, d, vocab));
  else
    KW_CUDA_OK(launch_pdl(PDL_EMBED, embed_kernel<float>, dim3(B), dim3(128), 0, st, tokens, ld_tokens, pos, (const float*)E, P, x, d, vocab));
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

int embed_seq(const int* tokens, int ld_tokens, int T, const void* E, const float* P, float* x, int B, int d, int vocab,
              kw_dtype t, cudaStream_t st) {
  if (t == KW_BF16) embed_seq_kernel<bf16><<<B * T, 128, 0, st>>>(tokens, ld_tokens, T, (const bf16*)E, P, x, d, vocab);
  else embed_seq_kernel<float><<<B * T, 128, 0, st>>>(tokens, ld_tokens, T, (const float*)E, P, x, d, vocab);
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

int convert_f32_to(const float* in, void* out, size_t n, kw_dtype t, cudaStream_t st) {
  int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 32);
  if (t == KW_BF16) convert_kernel<float, bf16><<<blocks, 256, 0, st>>>(in, (bf16*)out, n);
  else convert_kernel<float, float><<<blocks, 256, 0, st>>>(in, (float*)out, n);
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

}  // namespace kw
