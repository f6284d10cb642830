// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, SS operands) for N in {64,128,256}, issued by one thread from
// fixed shared-memory tiles, with / without a tcgen05.commit per group of 4.
#include <cstdio>
#include "../../kotoba_whisper_b200/csrc/tc_common.cuh"
namespace kw { void set_error(const char*, ...) {} }
using namespace kw::tc;
template <int N, int COMMIT>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 16384 + 32768, slot = bar + 64;
  uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)) + 16384 + 32768 + 64);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(bar + 8 * i, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) tmem_alloc(slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (warp == 0 && lane == 0) {
    const uint64_t da = make_desc(base), db = make_desc(base + 16384);
    constexpr uint32_t idesc = make_idesc(128, N, 0, 0);
    mbar_arrive(bar + 24);  // barrier 3: phase 0 completes now, so waiting on parity 0 always succeeds immediately
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (COMMIT >= 2) { mbar_wait(bar + 24, 0); tc_fence_after(); }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_f16(tmem, da + 2 * kk, db + 2 * kk, idesc, 1);
      if (COMMIT) umma_commit(bar + 8 * (it & 1));
    }
    long long t1 = clock64();
    umma_commit(bar + 16);
    mbar_wait(bar + 16, 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}
template <int N, int COMMIT> void run() {
  long long* out; cudaMalloc(&out, 16); int iters = 2000;
  cudaFuncSetAttribute(k<N, COMMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  k<N, COMMIT><<<1, 128, 60000>>>(out, iters); cudaDeviceSynchronize();
  k<N, COMMIT><<<1, 128, 60000>>>(out, iters); cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d commit/4=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (err %s)\n", N, COMMIT, (double)h[0] / (4.0 * iters), (double)h[1] / (4.0 * iters), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() { run<64, 0>(); run<64, 1>(); run<64, 2>(); run<128, 1>(); run<128, 2>(); run<256, 0>(); run<256, 1>(); run<256, 2>(); return 0; }
