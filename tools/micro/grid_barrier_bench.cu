// Micro-benchmark for round-2 planning: what does a grid-wide barrier between the phases of a persistent decode kernel
// cost on B200, compared with the ~2-5 us of a (PDL-chained) kernel boundary?  One CTA per SM, every CTA arrives on a global
// counter (red.release.gpu) and spins on it (ld.acquire.gpu); variants: 1 polling thread per CTA + bar.sync, or all
// threads polling; with and without a small amount of "work" (a dependent global write + read by the neighbour CTA).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

template <int WORK>
__global__ void k(unsigned* counter, float* buf, int rounds, unsigned long long* t_out) {
  unsigned long long t0 = 0;
  if (threadIdx.x == 0 && blockIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  float acc = 0.f;
  for (int r = 0; r < rounds; ++r) {
    if (WORK) {  // each CTA writes a line, the barrier makes it visible, the next CTA reads it
      buf[(size_t)blockIdx.x * 32 + (threadIdx.x & 31)] = acc + r;
    }
    grid_barrier(counter, (unsigned)(r + 1) * gridDim.x);
    if (WORK) acc += __ldcg(buf + (size_t)((blockIdx.x + 1) % gridDim.x) * 32 + (threadIdx.x & 31));
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    t_out[0] = t1 - t0;
  }
  if (acc == 12345.f) buf[0] = acc;
}

template <int WORK> void run(int ctas, int threads, const char* name) {
  unsigned* counter; float* buf; unsigned long long* t;
  cudaMalloc(&counter, 4); cudaMalloc(&buf, 1 << 20); cudaMalloc(&t, 8);
  const int rounds = 2000;
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(counter, 0, 4);
    void* args[] = {&counter, &buf, (void*)&rounds, &t};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k<WORK>, dim3(ctas), dim3(threads), args, 0, 0);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
    cudaDeviceSynchronize();
  }
  unsigned long long h; cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %3d CTAs x %4d thr: %7.2f us per barrier round\n", name, ctas, threads, h / 1e3 / rounds);
}

int main() {
  run<0>(148, 128, "barrier only");
  run<0>(148, 192, "barrier only");
  run<0>(120, 192, "barrier only");
  run<0>(40, 192, "barrier only");
  run<1>(148, 192, "barrier + line write / neighbour read");
  run<1>(296, 192, "barrier + write/read, 2 CTAs per SM");
  return 0;
}
