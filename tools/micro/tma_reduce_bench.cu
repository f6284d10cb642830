// Micro-benchmark: can the fp32 residual add of the encoder's out-projection / fc2 epilogues (x += tile, 491 MB per
// GEMM at B = 64) be handed to the L2 as a TMA reduction instead of a read-modify-write by the epilogue warps?
// Three ways to apply "x[i] += v" over a buffer far larger than L2, one CTA per SM, 8 warps:
//   rmw     ld.global.v4 + add + st.global.v4 (what the epilogue does today, here with everything in flight)
//   store   st.global.v4 only (upper bound of the SM -> memory path)
//   reduce  smem tile -> cp.reduce.async.bulk.global.shared::cta.add.f32 (16 KB per bulk op, two buffers per warp)
// Reported: GB/s of x traffic counted as read + write (2 x bytes) for rmw / reduce, 1 x bytes for store.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_rmw(float4* x, size_t n4, float v) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 4 * stride) {
    float4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * stride < n4) a[u] = x[i + u * stride];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * stride < n4) {
        a[u].x += v; a[u].y += v; a[u].z += v; a[u].w += v;
        x[i + u * stride] = a[u];
      }
  }
}

__global__ void __launch_bounds__(256) k_store(float4* x, size_t n4, float v) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const float4 a = make_float4(v, v, v, v);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) x[i] = a;
}

// each warp owns two CHUNK-byte shared buffers; it fills one (stand-in for TMEM -> registers -> st.shared), makes it
// visible to the async proxy and issues one bulk reduction; the buffer is reused when its bulk group has been read
template <int CHUNK>
__global__ void __launch_bounds__(256) k_reduce(float* x, size_t n, float v) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* buf = reinterpret_cast<float*>(smem) + (size_t)warp * 2 * (CHUNK / 4);
  const size_t per = CHUNK / 4, n_chunks = n / per, total_warps = (size_t)gridDim.x * 8;
  int which = 0;
  for (size_t c = (size_t)blockIdx.x * 8 + warp; c < n_chunks; c += total_warps, which ^= 1) {
    float* b = buf + which * per;
    // the buffer's previous bulk op (two iterations ago) must have finished READING shared memory
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    for (int i = lane * 4; i < (int)per; i += 128) *reinterpret_cast<float4*>(b + i) = make_float4(v, v, v, v);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      const unsigned s = (unsigned)__cvta_generic_to_shared(b);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(x + c * per), "r"(s),
                   "r"(CHUNK)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F> float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  const size_t n = (size_t)96000 * 1280;  // the encoder's residual stream at B = 64: 491.5 MB
  float* x;
  cudaMalloc(&x, n * 4);
  cudaMemset(x, 0, n * 4);
  int sm = 148;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  const double gb = n * 4 / 1e9;
  float ms = time_ms([&] { k_rmw<<<sm * 8, 256>>>((float4*)x, n / 4, 1.0f); }, 10);
  printf("rmw     (ld + add + st, 8 CTAs/SM)      %7.1f us  %7.0f GB/s (read + write)\n", ms * 1e3, 2 * gb / (ms / 1e3));
  ms = time_ms([&] { k_store<<<sm * 8, 256>>>((float4*)x, n / 4, 1.0f); }, 10);
  printf("store   (st only)                       %7.1f us  %7.0f GB/s (write)\n", ms * 1e3, gb / (ms / 1e3));
  cudaFuncSetAttribute(k_reduce<2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 2048);
  cudaFuncSetAttribute(k_reduce<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 8192);
  ms = time_ms([&] { k_reduce<2048><<<sm, 256, 8 * 2 * 2048>>>(x, n, 1.0f); }, 10);
  printf("reduce  (bulk add.f32, 2 KB ops, 1 CTA/SM) %7.1f us  %7.0f GB/s (read + write at the L2)\n", ms * 1e3, 2 * gb / (ms / 1e3));
  ms = time_ms([&] { k_reduce<8192><<<sm, 256, 8 * 2 * 8192>>>(x, n, 1.0f); }, 10);
  printf("reduce  (bulk add.f32, 8 KB ops, 1 CTA/SM) %7.1f us  %7.0f GB/s (read + write at the L2)\n", ms * 1e3, 2 * gb / (ms / 1e3));
  ms = time_ms([&] { k_reduce<2048><<<sm * 2, 256, 8 * 2 * 2048>>>(x, n, 1.0f); }, 10);
  printf("reduce  (bulk add.f32, 2 KB ops, 2 CTA/SM) %7.1f us  %7.0f GB/s (read + write at the L2)\n", ms * 1e3, 2 * gb / (ms / 1e3));
  // correctness: after all of the above x must hold a whole number of +1 steps everywhere
  float h[4];
  cudaMemcpy(h, x + n - 4, 16, cudaMemcpyDeviceToHost);
  printf("x[last] = %.1f  err=%s\n", h[3], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
