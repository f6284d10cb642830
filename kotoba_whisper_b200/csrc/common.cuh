// Shared helpers for the kwb200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/kwb200.h"

namespace kw {

typedef __nv_bfloat16 bf16;

void set_error(const char* fmt, ...);

#define KW_CUDA_OK(expr)                                                                      \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      kw::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));     \
      return KW_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define KW_LAUNCH_OK()                                                                        \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      kw::set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e));        \
      return KW_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define KW_REQUIRE(cond, ...)                                                                 \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      kw::set_error(__VA_ARGS__);                                                             \
      return KW_ERR_ARG;                                                                      \
    }                                                                                         \
  } while (0)

// ---- typed element access: storage type T in {float, bf16}, arithmetic always fp32 ---------------
__device__ __forceinline__ float ld_f(const float* p) { return *p; }
__device__ __forceinline__ float ld_f(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_f(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_f(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// 4 consecutive elements (pointer must be 16 B / 8 B aligned respectively)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

// exact GELU, 0.5 x (1 + erf(x / sqrt 2))  (HF/activations.py:70-89 -> torch gelu "none")
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Programmatic dependent launch: the kernel may start (and run up to its griddepcontrol.wait) while the preceding
// kernel in the stream is still draining.  Every kernel launched this way executes griddepcontrol.wait before it touches
// global memory, so the stream order of memory effects is preserved; the ~2-3 us launch latency of the short decode-step
// kernels overlaps the predecessor instead of adding to the per-position critical path.
enum PdlKind { PDL_LN = 1, PDL_SELF_ATTN = 2, PDL_CROSS_ATTN = 4, PDL_EMBED = 8, PDL_SAMPLE = 16 };
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(int kind, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  static const int mask = [] {
    const char* e = getenv("KW_PDL_MASK");  // bit set of PdlKind
    // Measured on B200, greedy pass of 124 positions at B = 64: none 60.5 ms, LN 57.5, LN + self-attention 55.2,
    // everything but cross-attention 56.3, everything 68.5 (an early-launched cross-attention grid — 1280 CTAs that need
    // every SM — ends up unevenly placed next to the still-resident GEMM CTAs).
    return e ? atoi(e) : (PDL_LN | PDL_SELF_ATTN);
  }();
  const int enabled = mask & kind;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// L2 eviction-priority policies for per-load cache hints
__device__ __forceinline__ uint64_t l2_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct SampleRules {
  int eos, pad, no_ts, ts_begin, max_initial, vocab;
};

// ---- GEMM front door (gemm_simt.cu / gemm_tc.cu) ---------------------------------------------------
enum Epilogue {
  EPI_STORE = 0,     // out = acc + bias
  EPI_GELU = 1,      // out = gelu(acc + bias)
  EPI_RESID = 2,     // out(f32) += acc + bias            (residual stream, in place)
  EPI_GELU_POS = 3,  // out(f32) = gelu(acc + bias) + pos[row % pos_period]   (conv2 + sinusoid table)
  EPI_ARGMAX = 4,    // decode-time vocabulary projection only: no logits leave the kernel — the three Whisper logits
                     // processors run on the accumulator and every epilogue warp emits arg-max partials (sample_combine)
};

// Arguments of the fused vocabulary-projection epilogue (EPI_ARGMAX) and of its combine kernel.  Vocabulary rows below
// tail0 (a multiple of 32, <= the first timestamp id: text ids only) are reduced inside the GEMM to one (best value, id)
// pair per batch row and 64-row half tile; rows >= tail0 (the ~1.5 k timestamp ids and the specials next to them) leave it as
// raw fp32 logits, because their rules need the whole timestamp range at once.
struct SampleFuse {
  const int* tokens;           // [B, ld_tokens] token history (columns <= pos are read)
  int ld_tokens, pos, begin_index, return_ts;
  const unsigned char* flags;  // [vocab] bit0 = suppress, bit1 = suppress at begin
  SampleRules rules;
  float2* vpart;               // [B][n_part] (best masked logit, its id as int bits) per 64-row half tile of text ids
  int n_part;                  // 2 * ceil(tail0 / 128)
  float* tail;                 // [B][tail_ld] raw logits of vocabulary rows tail0 .. vocab - 1
  int tail0, tail_ld;
};

// Optional tail of the combine kernel: embedding of the token just picked + the first LayerNorm of decoder layer 0 for the
// NEXT position (saves the embed and LayerNorm launches at the head of every decoder position of a greedy pass).
struct EmbedNext {
  const void* E;      // [vocab, d] bf16 token embedding
  const float* P;     // [max_target_pos, d] learned positions
  const float *ln_w, *ln_b;
  float* x;           // [B, d] residual stream of the next position
  void* da;           // [B, d] bf16 LayerNorm output (operand of the next position's QKV projection)
  int d, on;
};

struct GemmArgs {
  const void* A;      // [M, K] row-major, lda elements between rows
  const void* W;      // [N, K] row-major (nn.Linear layout)
  const float* bias;  // [N] or nullptr
  void* out;          // [M, N] row-major, ldo elements between rows
  const float* pos;   // EPI_GELU_POS only: [pos_period, N]
  int M, N, K, lda, ldo, pos_period;
  int epi;
  kw_dtype a_type, w_type, out_type;
  const SampleFuse* sample;  // EPI_ARGMAX only
  int w_hint;  // decode-time (skinny) launches: L2 eviction priority of the weight loads (0 normal, 1 evict_last, 2 evict_first)
};

int gemm_simt(const GemmArgs& g, cudaStream_t st);
// tcgen05 + TMA path (bf16 A, bf16 W); returns KW_ERR_UNSUPPORTED when the shape does not fit its tiling.
int gemm_tc(const GemmArgs& g, cudaStream_t st);
int gemm(const GemmArgs& g, cudaStream_t st);  // dispatch: tensor path when eligible and enabled

}  // namespace kw
