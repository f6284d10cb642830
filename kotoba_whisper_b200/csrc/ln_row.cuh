// LayerNorm of one row held in a warp's registers (lane owns float4 groups lane, lane + 32, ...): biased variance,
// two-pass, eps 1e-5 — shared by layernorm_kernel (elementwise.cu) and the embedding tail of sample_combine_kernel
// (sampling.cu), so that both produce the same bits.  HF/models/whisper/modeling_whisper.py:449-506 (nn.LayerNorm).
#pragma once
#include "common.cuh"

namespace kw {

template <typename T, int NV, bool EXACT>
__device__ __forceinline__ void ln_row(const float4 (&v)[NV], const float* __restrict__ w, const float* __restrict__ bias,
                                       T* __restrict__ orow, int d, int lane) {
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (EXACT || c < d) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(sum) / (float)d;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (EXACT || c < d) {
      float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, dd = v[i].w - mean;
      sq += (a * a + b * b) + (cc * cc + dd * dd);
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)d + 1e-5f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (EXACT || c < d) {
      float4 g = __ldg(reinterpret_cast<const float4*>(w + c)), be = __ldg(reinterpret_cast<const float4*>(bias + c));
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + be.x;
      o.y = (v[i].y - mean) * rstd * g.y + be.y;
      o.z = (v[i].z - mean) * rstd * g.z + be.z;
      o.w = (v[i].w - mean) * rstd * g.w + be.w;
      st4(orow + c, o);
    }
  }
}

}  // namespace kw
