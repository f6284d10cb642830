// Micro-benchmark: MUFU.EX2 throughput and the cost of the softmax inner loop body, per SM, vs resident warps.
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned pack(float a, float b) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<unsigned*>(&h); }
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float v[32];
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x * 0.001f + i * 0.01f;
  float rs0 = 0, rs1 = 0, rs2 = 0, rs3 = 0; unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float p0, p1, p2, p3;
      if (MODE == 0) { p0 = ex2(v[i]); p1 = ex2(v[i+1]); p2 = ex2(v[i+2]); p3 = ex2(v[i+3]); }
      else { p0 = ex2(fmaf(v[i], a, b)); p1 = ex2(fmaf(v[i+1], a, b)); p2 = ex2(fmaf(v[i+2], a, b)); p3 = ex2(fmaf(v[i+3], a, b)); }
      if (MODE >= 1) { rs0 += p0; rs1 += p1; rs2 += p2; rs3 += p3; }
      if (MODE >= 2) { acc ^= pack(p0, p1); acc ^= pack(p2, p3); }
      if (MODE == 0) { rs0 += p0 + p1 + p2 + p3; }
      v[i] = p0 * 1e-3f - 1.0f; v[i+1] = p1 * 1e-3f - 1.0f; v[i+2] = p2 * 1e-3f - 1.0f; v[i+3] = p3 * 1e-3f - 1.0f;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = rs0 + rs1 + rs2 + rs3 + acc;
}
template <int MODE> void run(int warps, const char* name) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  int iters = 2000; cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148, warps * 32>>>(out, 10, 1.44f, -3.f); cudaDeviceSynchronize();
  cudaEventRecord(a); k<MODE><<<148, warps * 32>>>(out, iters, 1.44f, -3.f); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double exps = 148.0 * warps * 32 * 32.0 * iters;
  printf("%-28s warps/SM %2d: %.3f ms  %.2f Gexp/s/SM  (%.2f exp/clk/SM @1.9GHz)\n", name, warps, ms, exps / ms / 1e6 / 148, exps / ms / 1e6 / 148 / 1.9);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 16, 32}) run<0>(w, "ex2 only");
  for (int w : {4, 8, 16, 32}) run<1>(w, "fma+ex2+sum (4 chains)");
  for (int w : {4, 8, 16, 32}) run<2>(w, "fma+ex2+sum+pack bf16");
  return 0;
}
