// tcgen05 / TMA tensor-core kernels (bf16 operands, fp32 accumulation in TMEM).  Placeholder until the kernels land:
// every entry reports KW_ERR_UNSUPPORTED so the dispatcher takes the SIMT path.
#include "common.cuh"

namespace kw {

int gemm_tc(const GemmArgs&, cudaStream_t) { return KW_ERR_UNSUPPORTED; }

int attention_tc(const void*, const void*, const void*, void*, int, int, int, int, long long, long long, long long,
                 long long, long long, long long, cudaStream_t) {
  return KW_ERR_UNSUPPORTED;
}

}  // namespace kw
