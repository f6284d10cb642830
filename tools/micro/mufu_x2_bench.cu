// Micro-benchmark: throughput of the packed half-precision exponentials (ex2.approx.ftz.bf16x2 / ex2.approx.f16x2) next to
// the fp32 MUFU.EX2 — does one packed instruction retire two exponentials per MUFU slot on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2_bf16x2(unsigned x) { unsigned y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned ex2_f16x2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
template <int MODE>
__global__ void k(unsigned* out, int iters) {
  unsigned v[16];
  float f[16];
  for (int i = 0; i < 16; ++i) { v[i] = 0xBF80BF80u ^ (threadIdx.x << 3) ^ i; f[i] = -1.0f - i * 0.01f - threadIdx.x * 1e-4f; }
  unsigned acc = 0; float facc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) { f[i] = ex2(f[i]) - 1.5f; }
      else if (MODE == 1) { v[i] = ex2_bf16x2(v[i]) ^ 0x80008000u; }
      else { v[i] = ex2_f16x2(v[i]) ^ 0x80008000u; }
    }
  }
  for (int i = 0; i < 16; ++i) { acc ^= v[i]; facc += f[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(facc);
}
template <int MODE> void run(int warps, const char* name, int per_instr) {
  unsigned* out; cudaMalloc(&out, 148 * 1024 * 4);
  int iters = 4000; cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148, warps * 32>>>(out, 10); cudaDeviceSynchronize();
  cudaEventRecord(a); k<MODE><<<148, warps * 32>>>(out, iters); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double exps = 148.0 * warps * 32 * 16.0 * iters * per_instr;
  printf("%-24s warps/SM %2d: %.3f ms  %.2f exp/clk/SM @1.9GHz (%d per instruction)\n", name, warps, ms, exps / ms / 1e6 / 148 / 1.9, per_instr);
  cudaFree(out);
}
int main() {
  for (int w : {8, 16, 32}) run<0>(w, "ex2.approx.ftz.f32", 1);
  for (int w : {8, 16, 32}) run<1>(w, "ex2.approx.ftz.bf16x2", 2);
  for (int w : {8, 16, 32}) run<2>(w, "ex2.approx.f16x2", 2);
  return 0;
}
