// The model handle behind the C ABI (include/kwb200.h): configuration, weight table and the device pools.  Shared by
// api.cu (kernel schedules) and decode_fused.cu (the persistent decode kernel's host side).
#pragma once
#include <vector>

#include "common.cuh"

namespace kw {
struct FusedDecode;  // decode_fused.cu
}

struct kw_model {
  kw_config cfg;
  kw_weights w;
  std::vector<kw_enc_layer_weights> enc;
  std::vector<kw_dec_layer_weights> dec;
  kw::SampleRules rules;
  kw_dtype t;
  // device pools
  char* pool = nullptr;
  size_t pool_bytes = 0;
  unsigned char* flags = nullptr;  // [vocab] bit0 = suppress, bit1 = suppress at begin
  // encoder workspaces
  void *bufP = nullptr, *bufQ = nullptr, *a = nullptr, *o = nullptr, *enc_out = nullptr;
  float* x = nullptr;
  // decoder workspaces
  float *dx = nullptr, *dqkv = nullptr, *dq = nullptr, *logits = nullptr;
  void *da = nullptr, *dattn = nullptr, *dh = nullptr;  // projection operands: model dtype (bf16 feeds the tcgen05 path)
  void* self_k = nullptr;  // [L][B][H][max_t][64]
  void* self_v = nullptr;
  void* xkv = nullptr;  // [L][B*S][2d]
  int* finished = nullptr;
  float* vpart = nullptr;  // [max_batch][2 * text tiles] float2 arg-max partials, one per 64-row half tile of text ids (EPI_ARGMAX)
  int* finished_host = nullptr;  // pinned
  cudaEvent_t finished_copied = nullptr;
  int enc_B = 0;
  // CUDA-graph replay of the decoder positions (api.cu kw_greedy_pass): captured chunks, capture stream, token buffer
  struct PassGraph {
    int B, n_prompt, max_length, return_ts, pos0, pos1, pre_in, cfg_epoch;
    cudaGraphExec_t exec;
    int n_launch;
    bool pre_out;
  };
  std::vector<PassGraph> graphs;
  cudaStream_t cap_stream = nullptr;
  int* gtokens = nullptr;  // [max_batch, max_target_pos] tokens of the pass being decoded (graphs bake this pointer)
  kw::FusedDecode* fused = nullptr;  // persistent decode kernel state (bf16 models), built lazily
};

