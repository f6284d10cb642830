// Micro-benchmark: DFMA / DADD / DMUL throughput per SM on B200 (the log-mel FFT runs in float64; its compute bound is
// (fp64 operations per frame) / (this rate)).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters) {
  double v[8];
  for (int i = 0; i < 8; ++i) v[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double a = 1.0000001, b = 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) v[i] = fma(v[i], a, b);
      else if (MODE == 1) v[i] = v[i] + b;
      else v[i] = v[i] * a;
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(int warps, const char* name) {
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  const int iters = 4000; cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148, warps * 32>>>(out, 10); cudaDeviceSynchronize();
  cudaEventRecord(a); k<MODE><<<148, warps * 32>>>(out, iters); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double ops = 148.0 * warps * 32 * 8.0 * iters;
  printf("%-6s warps/SM %2d: %.3f ms  %.2f ops/clk/SM @1.9GHz  (%.2f Tops/s; x2 flops for fma)\n", name, warps, ms,
         ops / ms / 1e6 / 148 / 1.9, ops / ms / 1e9);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 16, 32}) run<0>(w, "dfma");
  for (int w : {8, 32}) run<1>(w, "dadd");
  for (int w : {8, 32}) run<2>(w, "dmul");
  return 0;
}
