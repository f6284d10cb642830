// C ABI of libkwb200.so (include/kwb200.h): model handle, workspace pools and the kernel schedules for the encoder,
// the one-shot cross-K/V projection, a decoder step and a whole greedy pass.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <vector>

#include "common.cuh"

namespace kw {

std::atomic<long long> g_launches{0};
static thread_local char g_err[1024] = "";
static std::atomic<int> g_gemm_impl{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// kernels defined in the other translation units
#ifdef KW_LOGMEL_TIMING
void logmel_debug_dump();
#endif
int logmel_launch(const float* audio, const long long* starts, const int32_t* lens, int B, int n_samples, int n_mels,
                  float* out, float* clip_max, cudaStream_t st);
void mel_filterbank_f64(int n_mels, std::vector<double>& fb);
int im2col_conv1(const float* mel, void* A1, int B, int C, int Tn, kw_dtype t, cudaStream_t st);
int im2col_conv2(const void* h0, void* A2, int B, int d, int Tin, int Tout, kw_dtype t, cudaStream_t st);
int layernorm(const float* x, const float* w, const float* b, void* out, int rows, int d, kw_dtype t, cudaStream_t st);
int embed(const int* tokens, int ld_tokens, int pos, const void* E, const float* P, float* x, int B, int d, int vocab,
          kw_dtype t, cudaStream_t st);
int convert_f32_to(const float* in, void* out, size_t n, kw_dtype t, cudaStream_t st);
int attention_simt(const void* q, const void* k, const void* v, void* out, int B, int H, int Tq, int Tk,
                   long long q_sb, long long q_st, long long kv_sb, long long kv_st, long long o_sb, long long o_st,
                   kw_dtype t, cudaStream_t st, int causal = 0);
int embed_seq(const int* tokens, int ld_tokens, int T, const void* E, const float* P, float* x, int B, int d, int vocab,
              kw_dtype t, cudaStream_t st);
int attention_tc(const void* q, const void* k, const void* v, void* out, int B, int H, int Tq, int Tk, long long q_sb,
                 long long q_st, long long kv_sb, long long kv_st, long long o_sb, long long o_st, cudaStream_t st);
void attention_tc_debug(int lbo, int sbo, int kstep);
void gemm_tc_set_stamps(unsigned long long* p);
void gemm_tc_set_2cta(int on);
int dec_self_attn(const float* qkv, void* kc, void* vc, void* out, int B, int d, int H, int max_t, int pos, kw_dtype t,
                  cudaStream_t st);
int dec_cross_attn(const float* q, const void* xkv, void* out, int B, int d, int H, int S, kw_dtype t, cudaStream_t st,
                   int hint, int head_rows);
int sample_launch(const float* logits, const unsigned char* flags, const SampleRules& r, int* tokens, int ld_tokens,
                  int B, int pos, int begin_index, int return_ts, int* finished, cudaStream_t st);

// decode_fused.cu: the persistent decode kernel (bf16 models)
int fused_decode_prepare(kw_model* m);
int fused_decode_pass(kw_model* m, int B, int n_prompt, int max_length, int return_ts, int* tokens, cudaStream_t st);
void fused_decode_destroy(kw_model* m);
// 0 = one kernel per op (default: measured faster), 1 = persistent fused kernel when the model fits,
// 2 = fused kernel required (error instead of falling back)
static std::atomic<int> g_decode_impl{[] {
  const char* e = getenv("KW_DECODE_FUSED");
  return e ? atoi(e) : 0;
}()};

int sample_combine_launch(const SampleFuse& sf, int* tokens, int B, int* finished, const EmbedNext& en, cudaStream_t st);
// 1 (default): the decode-time vocabulary projection of a bf16 model runs the logits processors + arg-max in its epilogue
// (EPI_ARGMAX) instead of writing fp32 logits for sample_kernel; 0: separate kernels (KW_SAMPLE_FUSED=0, A/B reference)
static std::atomic<int> g_sample_fused{[] {
  const char* e = getenv("KW_SAMPLE_FUSED");
  return e ? atoi(e) : 1;
}()};

// 1 (default): kw_greedy_pass replays captured CUDA graphs of the decoder positions (KW_DECODE_GRAPH=0: eager launches)
static std::atomic<int> g_decode_graph{[] {
  const char* e = getenv("KW_DECODE_GRAPH");
  return e ? atoi(e) : 1;
}()};

int gemm(const GemmArgs& g, cudaStream_t st) {
  const int impl = g_gemm_impl.load();
  if (impl != 1 && g.a_type == KW_BF16 && g.w_type == KW_BF16) {
    int rc = gemm_tc(g, st);
    if (rc != KW_ERR_UNSUPPORTED) return rc;
    if (impl == 2) return rc;
  } else if (impl == 2) {
    set_error("gemm: tcgen05 path forced but operands are not bf16");
    return KW_ERR_UNSUPPORTED;
  }
  return gemm_simt(g, st);
}

static size_t esize(kw_dtype t) { return t == KW_BF16 ? 2 : 4; }

// The kernels keep per-process state that belongs to ONE device (the > 48 KB shared-memory opt-ins, the SM count, the
// log-mel tables): the library is one-process-per-GPU by design.  Every entry point that launches work
// checks that the current device is the one the process first used, instead of failing obscurely on a second GPU.
static std::atomic<int> g_bound_device{-1};
static int bound_device_ok(const char* who) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return KW_OK;  // no device at all: argument checks first, the first CUDA call then fails loudly (no CPU fallback)
  }
  int expect = -1;
  if (g_bound_device.compare_exchange_strong(expect, dev) || expect == dev) return KW_OK;
  set_error("%s: this process is bound to CUDA device %d (first use) but device %d is current; run one process per GPU",
            who, expect, dev);
  return KW_ERR_ARG;
}


// ---- optional per-category device timing (CUDA events on the launching stream; bench.py's roofline legs) -----------
struct ProfRec {
  int cat;
  cudaEvent_t a, b;
};
struct Profiler {
  unsigned mask = 0;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  double work[KW_PROF_NCAT] = {0};
  long long count[KW_PROF_NCAT] = {0};
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
};
static Profiler g_prof;

struct ProfScope {
  int cat;
  cudaStream_t st;
  cudaEvent_t b = nullptr;
  ProfScope(int cat_, double work, cudaStream_t st_) : cat(cat_), st(st_) {
    if (!(g_prof.mask & (1u << cat))) return;
    cudaEvent_t a = g_prof.get();
    b = g_prof.get();
    cudaEventRecord(a, st);
    g_prof.recs.push_back({cat, a, b});
    g_prof.work[cat] += work;
    g_prof.count[cat] += 1;
  }
  ~ProfScope() {
    if (b) cudaEventRecord(b, st);
  }
};
static double gemm_flops(const GemmArgs& g) { return 2.0 * g.M * (double)g.N * g.K; }

}  // namespace kw

using namespace kw;

#include "model.cuh"

#define KW_TRY(expr)           \
  do {                         \
    int _rc = (expr);          \
    if (_rc != KW_OK) return _rc; \
  } while (0)

extern "C" {

const char* kw_last_error(void) { return g_err; }
const char* kw_version(void) { return "kwb200 0.1 (sm_100a)"; }
void kw_set_gemm_impl(int32_t impl) { g_gemm_impl.store(impl); }
void kw_set_gemm_2cta(int32_t on) { gemm_tc_set_2cta(on); }
void kw_set_decode_impl(int32_t impl) { g_decode_impl.store(impl); }
void kw_set_sample_fused(int32_t on) { g_sample_fused.store(on); }
void kw_set_decode_graph(int32_t on) { g_decode_graph.store(on); }

void kw_debug_attention_desc(int32_t v_lbo_bytes, int32_t v_sbo_bytes, int32_t v_kstep_bytes) {
  attention_tc_debug(v_lbo_bytes, v_sbo_bytes, v_kstep_bytes);
}

void kw_debug_gemm_stamps(uint64_t* dev_buffer_16) { gemm_tc_set_stamps((unsigned long long*)dev_buffer_16); }

void kw_profile_enable(uint32_t category_mask) { g_prof.mask = category_mask; }

int kw_profile_read(int32_t category, double* total_ms, int64_t* launches, double* work, int32_t reset) {
  KW_REQUIRE(category >= 0 && category < KW_PROF_NCAT, "kw_profile_read: bad category %d", category);
  double ms = 0.0;
  for (auto& r : g_prof.recs) {
    if (r.cat != category) continue;
    float t = 0.0f;
    KW_CUDA_OK(cudaEventSynchronize(r.b));
    KW_CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t;
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = g_prof.count[category];
  if (work) *work = g_prof.work[category];
  if (reset) {
    std::vector<ProfRec> keep;
    for (auto& r : g_prof.recs) {
      if (r.cat == category) {
        g_prof.pool.push_back(r.a);
        g_prof.pool.push_back(r.b);
      } else {
        keep.push_back(r);
      }
    }
    g_prof.recs.swap(keep);
    g_prof.work[category] = 0;
    g_prof.count[category] = 0;
  }
  return KW_OK;
}
int64_t kw_launch_count(int32_t reset) {
  long long v = g_launches.load();
  if (reset) g_launches.store(0);
  return v;
}

int kw_logmel(const float* audio, const int32_t* lens, int32_t B, int32_t n_samples, int32_t n_mels, float* out,
              float* clip_max, kw_stream stream) {
  KW_TRY(bound_device_ok("kw_logmel"));
  KW_REQUIRE(audio && out && clip_max, "kw_logmel: null pointer");
  ProfScope ps(KW_PROF_LOGMEL, (double)B * (4.0 * n_samples + 4.0 * n_mels * (n_samples / 160)), (cudaStream_t)stream);
  return logmel_launch(audio, nullptr, lens, B, n_samples, n_mels, out, clip_max, (cudaStream_t)stream);
}

int kw_logmel_windows(const float* recording, const int64_t* starts, const int32_t* lens, int32_t W, int32_t n_samples,
                      int32_t n_mels, float* out, float* clip_max, kw_stream stream) {
  KW_TRY(bound_device_ok("kw_logmel_windows"));
  KW_REQUIRE(recording && starts && lens && out && clip_max, "kw_logmel_windows: null pointer");
  ProfScope ps(KW_PROF_LOGMEL, (double)W * (4.0 * n_samples + 4.0 * n_mels * (n_samples / 160)), (cudaStream_t)stream);
  return logmel_launch(recording, (const long long*)starts, lens, W, n_samples, n_mels, out, clip_max,
                       (cudaStream_t)stream);
}

int kw_mel_filterbank(int32_t n_mels, double* out_host) {
#ifdef KW_LOGMEL_TIMING
  if (n_mels < 0) {  // debug build: dump + reset the log-mel phase counters
    kw::logmel_debug_dump();
    return KW_OK;
  }
#endif
  KW_REQUIRE(n_mels > 0 && n_mels <= 128 && out_host, "kw_mel_filterbank: bad arguments");
  std::vector<double> fb;
  mel_filterbank_f64(n_mels, fb);
  memcpy(out_host, fb.data(), fb.size() * sizeof(double));
  return KW_OK;
}

int kw_model_create(const kw_config* cfg, const kw_weights* w, const kw_token_rules* rules, kw_model** out) {
  KW_TRY(bound_device_ok("kw_model_create"));
  KW_REQUIRE(cfg && w && rules && out, "kw_model_create: null argument");
  KW_REQUIRE(cfg->d_model % 64 == 0 && cfg->n_heads * 64 == cfg->d_model, "kw_model_create: head dim must be 64");
  KW_REQUIRE(cfg->ffn_dim % 16 == 0 && (3 * cfg->n_mels) % 16 == 0, "kw_model_create: ffn / 3*n_mels must be multiples of 16");
  KW_REQUIRE(cfg->max_batch >= 1 && cfg->max_target_pos <= 512, "kw_model_create: bad max_batch / max_target_pos");
  KW_REQUIRE(cfg->dtype == KW_F32 || cfg->dtype == KW_BF16, "kw_model_create: bad dtype");
  kw_model* m = new kw_model();
  m->cfg = *cfg;
  m->w = *w;
  m->enc.assign(w->enc, w->enc + cfg->enc_layers);
  m->dec.assign(w->dec, w->dec + cfg->dec_layers);
  m->w.enc = m->enc.data();
  m->w.dec = m->dec.data();
  m->t = (kw_dtype)cfg->dtype;
  m->rules = {rules->eos_token_id, rules->pad_token_id, rules->no_timestamps_token_id,
              rules->no_timestamps_token_id + 1, rules->max_initial_timestamp_index, cfg->vocab_size};

  const size_t B = cfg->max_batch, d = cfg->d_model, S = cfg->max_source_pos, T2 = 2 * S, F = cfg->ffn_dim;
  const size_t V = cfg->vocab_size, L = cfg->dec_layers, es = esize(m->t);
  auto al = [](size_t n) { return (n + 255) / 256 * 256; };
  const size_t szP = al(std::max(std::max(B * T2 * 3 * cfg->n_mels, B * S * 3 * d), B * S * 3 * d) * es);
  const size_t szQ = al(std::max(B * T2 * d, B * S * F) * es);
  const size_t szA = al(B * S * d * es), szX = al(B * S * d * 4);
  const size_t szSelf = al(L * B * d * cfg->max_target_pos * es), szXkv = al(L * B * S * 2 * d * es);
  const size_t szDec = al(B * d * 4), szDqkv = al(B * 3 * d * 4), szDh = al(B * F * 4), szLog = al(B * (V + 32) * 4);
  const size_t szVp = al(B * (V / 32 + 1) * sizeof(float2));  // arg-max partials of the fused vocabulary epilogue
  const size_t szTok = al(B * cfg->max_target_pos * sizeof(int));
  const size_t total = szP + szQ + 3 * szA + szX + 2 * szSelf + szXkv + 4 * szDec + szDqkv + szDh + szLog + al(V) +
                       al(B * 4) + szVp + szTok;
  cudaError_t e = cudaMalloc(&m->pool, total);
  if (e != cudaSuccess) {
    set_error("kw_model_create: cudaMalloc(%zu bytes) failed: %s", total, cudaGetErrorString(e));
    delete m;
    return KW_ERR_NOMEM;
  }
  m->pool_bytes = total;
  char* p = m->pool;
  auto take = [&](size_t n) { char* r = p; p += n; return (void*)r; };
  m->bufP = take(szP); m->bufQ = take(szQ);
  m->a = take(szA); m->o = take(szA); m->enc_out = take(szA);
  m->x = (float*)take(szX);
  m->self_k = take(szSelf); m->self_v = take(szSelf); m->xkv = take(szXkv);
  m->dx = (float*)take(szDec); m->da = take(szDec); m->dq = (float*)take(szDec); m->dattn = take(szDec);
  m->dqkv = (float*)take(szDqkv); m->dh = take(szDh); m->logits = (float*)take(szLog);
  m->flags = (unsigned char*)take(al(V));
  m->finished = (int*)take(al(B * 4));
  m->vpart = (float*)take(szVp);
  m->gtokens = (int*)take(szTok);

  std::vector<unsigned char> hf(V, 0);
  for (int i = 0; i < rules->n_suppress; ++i) {
    int t = rules->suppress_tokens[i];
    if (t >= 0 && (size_t)t < V) hf[t] |= 1;
  }
  for (int i = 0; i < rules->n_begin_suppress; ++i) {
    int t = rules->begin_suppress_tokens[i];
    if (t >= 0 && (size_t)t < V) hf[t] |= 2;
  }
  if (cudaMemcpy(m->flags, hf.data(), V, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMallocHost(&m->finished_host, B * sizeof(int)) != cudaSuccess) {
    set_error("kw_model_create: flag upload / pinned alloc failed");
    cudaFree(m->pool);
    delete m;
    return KW_ERR_CUDA;
  }
  *out = m;
  return KW_OK;
}

void kw_model_destroy(kw_model* m) {
  if (!m) return;
  fused_decode_destroy(m);
  for (auto& g : m->graphs) cudaGraphExecDestroy(g.exec);
  if (m->cap_stream) cudaStreamDestroy(m->cap_stream);
  cudaFree(m->pool);
  if (m->finished_host) cudaFreeHost(m->finished_host);
  if (m->finished_copied) cudaEventDestroy(m->finished_copied);
  delete m;
}

int64_t kw_model_workspace_bytes(const kw_model* m) { return m ? (int64_t)m->pool_bytes : 0; }

static GemmArgs mk(const void* A, int lda, kw_dtype at, const void* W, kw_dtype wt, const float* bias, void* out,
                   int ldo, kw_dtype ot, int M, int N, int K, int epi) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = A; g.W = W; g.bias = bias; g.out = out;
  g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldo = ldo;
  g.epi = epi; g.a_type = at; g.w_type = wt; g.out_type = ot;
  return g;
}

static int gemm_p(int cat, const GemmArgs& g, cudaStream_t st) {
  ProfScope ps(cat, gemm_flops(g), st);
  return gemm(g, st);
}

// The encoder in three pieces so that kw_encode_decode can slot decoder positions of ANOTHER batch between layer groups.
static int encode_stem(kw_model* m, const float* mel, int B, cudaStream_t st) {
  const kw_config& c = m->cfg;
  const kw_dtype t = m->t;
  const int d = c.d_model, S = c.max_source_pos, T2 = 2 * S, M = B * S;
  // conv stem as two im2col GEMMs (modeling_whisper.py:619-625)
  KW_TRY(im2col_conv1(mel, m->bufP, B, c.n_mels, T2, t, st));
  KW_TRY(gemm_p(KW_PROF_ENC_GEMM, mk(m->bufP, 3 * c.n_mels, t, m->w.conv1_w, t, m->w.conv1_b, m->bufQ, d, t, B * T2, d, 3 * c.n_mels, EPI_GELU), st));
  KW_TRY(im2col_conv2(m->bufQ, m->bufP, B, d, T2, S, t, st));
  GemmArgs g = mk(m->bufP, 3 * d, t, m->w.conv2_w, t, m->w.conv2_b, m->x, d, KW_F32, M, d, 3 * d, EPI_GELU_POS);
  g.pos = m->w.enc_pos;
  g.pos_period = S;
  return gemm_p(KW_PROF_ENC_GEMM, g, st);
}

static int encode_layers(kw_model* m, int B, int l0, int l1, cudaStream_t st) {
  const kw_config& c = m->cfg;
  const kw_dtype t = m->t;
  const int d = c.d_model, S = c.max_source_pos, F = c.ffn_dim, M = B * S;
  for (int l = l0; l < l1 && l < c.enc_layers; ++l) {
    const kw_enc_layer_weights& w = m->enc[l];
    KW_TRY(layernorm(m->x, w.ln1_w, w.ln1_b, m->a, M, d, t, st));
    KW_TRY(gemm_p(KW_PROF_ENC_GEMM, mk(m->a, d, t, w.wqkv, t, w.bqkv, m->bufP, 3 * d, t, M, 3 * d, d, EPI_STORE), st));
    {
      const char* qkv = (const char*)m->bufP;
      const size_t es = esize(t);
      ProfScope ps(KW_PROF_ENC_ATTN, 4.0 * B * c.n_heads * (double)S * S * 64, st);
      int rc = KW_ERR_UNSUPPORTED;
      if (t == KW_BF16 && g_gemm_impl.load() != 1)
        rc = attention_tc(qkv, qkv + d * es, qkv + 2 * d * es, m->o, B, c.n_heads, S, S, (long long)S * 3 * d, 3 * d,
                          (long long)S * 3 * d, 3 * d, (long long)S * d, d, st);
      if (rc == KW_ERR_UNSUPPORTED)
        rc = attention_simt(qkv, qkv + d * es, qkv + 2 * d * es, m->o, B, c.n_heads, S, S, (long long)S * 3 * d, 3 * d,
                            (long long)S * 3 * d, 3 * d, (long long)S * d, d, t, st);
      KW_TRY(rc);
    }
    KW_TRY(gemm_p(KW_PROF_ENC_GEMM, mk(m->o, d, t, w.wo, t, w.bo, m->x, d, KW_F32, M, d, d, EPI_RESID), st));
    KW_TRY(layernorm(m->x, w.ln2_w, w.ln2_b, m->a, M, d, t, st));
    KW_TRY(gemm_p(KW_PROF_ENC_GEMM, mk(m->a, d, t, w.w1, t, w.b1, m->bufQ, F, t, M, F, d, EPI_GELU), st));
    KW_TRY(gemm_p(KW_PROF_ENC_GEMM, mk(m->bufQ, F, t, w.w2, t, w.b2, m->x, d, KW_F32, M, d, F, EPI_RESID), st));
  }
  return KW_OK;
}

static int encode_finish(kw_model* m, int B, float* enc_out, cudaStream_t st) {
  const kw_config& c = m->cfg;
  const int M = B * c.max_source_pos, d = c.d_model;
  KW_TRY(layernorm(m->x, m->w.enc_ln_w, m->w.enc_ln_b, m->enc_out, M, d, m->t, st));
  if (enc_out) KW_TRY(layernorm(m->x, m->w.enc_ln_w, m->w.enc_ln_b, enc_out, M, d, KW_F32, st));
  m->enc_B = B;
  return KW_OK;
}

int kw_encode(kw_model* m, const float* mel, int32_t B, float* enc_out, kw_stream stream) {
  KW_TRY(bound_device_ok("kw_encode"));
  KW_REQUIRE(m && mel, "kw_encode: null argument");
  KW_REQUIRE(B >= 1 && B <= m->cfg.max_batch, "kw_encode: B=%d outside [1, max_batch=%d]", B, m->cfg.max_batch);
  cudaStream_t st = (cudaStream_t)stream;
  KW_TRY(encode_stem(m, mel, B, st));
  KW_TRY(encode_layers(m, B, 0, m->cfg.enc_layers, st));
  return encode_finish(m, B, enc_out, st);
}

int kw_set_encoder_output(kw_model* m, const float* enc, int32_t B, kw_stream stream) {
  KW_TRY(bound_device_ok("kw_set_encoder_output"));
  KW_REQUIRE(m && enc && B >= 1 && B <= m->cfg.max_batch, "kw_set_encoder_output: bad arguments");
  KW_TRY(convert_f32_to(enc, m->enc_out, (size_t)B * m->cfg.max_source_pos * m->cfg.d_model, m->t, (cudaStream_t)stream));
  m->enc_B = B;
  return KW_OK;
}

int kw_cross_kv(kw_model* m, int32_t B, kw_stream stream) {
  KW_TRY(bound_device_ok("kw_cross_kv"));
  KW_REQUIRE(m && B >= 1 && B <= m->enc_B, "kw_cross_kv: B=%d but encoder output holds %d rows", B, m ? m->enc_B : 0);
  const kw_config& c = m->cfg;
  const int d = c.d_model, S = c.max_source_pos;
  const size_t layer_stride = (size_t)c.max_batch * S * 2 * d * esize(m->t);
  for (int l = 0; l < c.dec_layers; ++l) {
    const kw_dec_layer_weights& w = m->dec[l];
    KW_TRY(gemm_p(KW_PROF_XKV_GEMM, mk(m->enc_out, d, m->t, w.wkv_x, m->t, w.bkv_x, (char*)m->xkv + l * layer_stride, 2 * d, m->t, B * S, 2 * d,
                   d, EPI_STORE), (cudaStream_t)stream));
  }
  return KW_OK;
}

// pre_embedded: the residual stream of this position and LayerNorm_1 of layer 0 were already written by the preceding
// position's sample_combine_kernel (EmbedNext)
static int decode_hidden(kw_model* m, const int32_t* tokens, int ld_tokens, int B, int pos, cudaStream_t st,
                         bool pre_embedded = false) {
  const kw_config& c = m->cfg;
  const kw_dtype t = m->t;
  const int d = c.d_model, S = c.max_source_pos, F = c.ffn_dim, H = c.n_heads, MT = c.max_target_pos;
  const size_t self_stride = (size_t)c.max_batch * d * MT * esize(t);
  const size_t xkv_stride = (size_t)c.max_batch * S * 2 * d * esize(t);
  // bring-up / attribution only (KW_DECODE_SKIP, bit set of kernel kinds to leave out of the chain; results are garbage,
  // the time difference is that kind's cost inside the dependent chain): 1 LayerNorm, 2 qkv, 4 self-attention, 8 out,
  // 16 cross-q, 32 cross-attention, 64 cross-out, 128 fc1, 256 fc2, 1024 embedding  (512 = vocabulary + sampling, below)
  static const int skip = getenv("KW_DECODE_SKIP") ? atoi(getenv("KW_DECODE_SKIP")) : 0;
  if (!pre_embedded && !(skip & 1024)) KW_TRY(embed(tokens, ld_tokens, pos, m->w.tok_embed, m->w.dec_pos, m->dx, B, d, c.vocab_size, t, st));
  // L2 eviction priorities of the decode step (measured on B200, greedy pass alone at B = 64, same box): cross-attention
  // K/V stream evict_first + decoder-layer weights evict_last + vocabulary matrix evict_first 45.97 -> 45.46 ms.  Keeping
  // a head of the K/V rows resident (evict_last on 64-192 rows per utterance: 45.9-46.1 ms) or pulling them into L2 from
  // the projections that run before each cross-attention (cp.async.bulk.prefetch.L2, 63-94 MB per cross-attention:
  // 46.3-47.0 ms) did not pay: the projections are latency-bound and slow down under the extra traffic by as much as
  // the cross-attention gains.  KW_XA_HINT / KW_W_HINT = 0 turn the hints off.
  static const int xa_hint = getenv("KW_XA_HINT") ? atoi(getenv("KW_XA_HINT")) : 1;
  static const int w_hint = getenv("KW_W_HINT") ? atoi(getenv("KW_W_HINT")) : 1;
  auto with_l2 = [&](GemmArgs g) {
    g.w_hint = w_hint;
    return g;
  };
  for (int l = 0; l < c.dec_layers; ++l) {
    const kw_dec_layer_weights& w = m->dec[l];
    if (!(pre_embedded && l == 0) && !(skip & 1)) KW_TRY(layernorm(m->dx, w.ln1_w, w.ln1_b, m->da, B, d, t, st));
    if (!(skip & 2))
      KW_TRY(gemm_p(KW_PROF_DEC_GEMM, with_l2(mk(m->da, d, t, w.wqkv, t, w.bqkv, m->dqkv, 3 * d, KW_F32, B, 3 * d, d, EPI_STORE)), st));
    if (!(skip & 4))
      KW_TRY(dec_self_attn(m->dqkv, (char*)m->self_k + l * self_stride, (char*)m->self_v + l * self_stride, m->dattn, B, d,
                           H, MT, pos, t, st));
    if (!(skip & 8))
      KW_TRY(gemm_p(KW_PROF_DEC_GEMM, with_l2(mk(m->dattn, d, t, w.wo, t, w.bo, m->dx, d, KW_F32, B, d, d, EPI_RESID)), st));
    if (!(skip & 1)) KW_TRY(layernorm(m->dx, w.lnx_w, w.lnx_b, m->da, B, d, t, st));
    if (!(skip & 16))
      KW_TRY(gemm_p(KW_PROF_DEC_GEMM, with_l2(mk(m->da, d, t, w.wq_x, t, w.bq_x, m->dq, d, KW_F32, B, d, d, EPI_STORE)), st));
    if (!(skip & 32)) {
      ProfScope ps(KW_PROF_DEC_CROSS, (double)B * S * 2 * d * esize(t), st);  // algorithmic bytes: K and V read once
      KW_TRY(dec_cross_attn(m->dq, (char*)m->xkv + l * xkv_stride, m->dattn, B, d, H, S, t, st, xa_hint, 0));
    }
    if (!(skip & 64))
      KW_TRY(gemm_p(KW_PROF_DEC_GEMM, with_l2(mk(m->dattn, d, t, w.wo_x, t, w.bo_x, m->dx, d, KW_F32, B, d, d, EPI_RESID)), st));
    if (!(skip & 1)) KW_TRY(layernorm(m->dx, w.ln3_w, w.ln3_b, m->da, B, d, t, st));
    if (!(skip & 128)) KW_TRY(gemm_p(KW_PROF_DEC_GEMM, with_l2(mk(m->da, d, t, w.w1, t, w.b1, m->dh, F, t, B, F, d, EPI_GELU)), st));
    if (!(skip & 256)) KW_TRY(gemm_p(KW_PROF_DEC_GEMM, with_l2(mk(m->dh, F, t, w.w2, t, w.b2, m->dx, d, KW_F32, B, d, F, EPI_RESID)), st));
  }
  return KW_OK;
}

// One decoder position.  pre_embedded: see decode_hidden.  want_next: when the fused vocabulary epilogue runs, let its combine
// kernel also embed the picked token and apply LayerNorm_1 of layer 0 for position pos + 1; *did_next reports whether
// that happened (the caller then passes pre_embedded for the next position).
static int decode_step(kw_model* m, int32_t* tokens, int32_t ld_tokens, int32_t B, int32_t pos, int32_t begin_index,
                       int32_t sample, int32_t return_timestamps, int32_t* finished, float* logits_out, cudaStream_t st,
                       bool pre_embedded, bool want_next, bool* did_next) {
  const kw_config& c = m->cfg;
  if (did_next) *did_next = false;
  KW_TRY(decode_hidden(m, tokens, ld_tokens, B, pos, st, pre_embedded));
  if (!sample && !logits_out) return KW_OK;
  static const int skip = getenv("KW_DECODE_SKIP") ? atoi(getenv("KW_DECODE_SKIP")) : 0;
  if (skip & 512) return KW_OK;
  KW_TRY(layernorm(m->dx, m->w.dec_ln_w, m->w.dec_ln_b, m->da, B, c.d_model, m->t, st));
  float* lg = logits_out ? logits_out : m->logits;
  if (sample && !logits_out && m->t == KW_BF16 && g_sample_fused.load() && g_gemm_impl.load() != 1) {
    // vocabulary projection with the logits processors + arg-max fused into its epilogue: no logits are materialised
    // (only the ~1.5 k timestamp-range logits per row, in the otherwise unused logits buffer)
    KW_REQUIRE(finished, "kw_decode_step: sample requires the finished array");
    SampleFuse sf;
    sf.tokens = tokens; sf.ld_tokens = ld_tokens; sf.pos = pos; sf.begin_index = begin_index;
    sf.return_ts = return_timestamps; sf.flags = m->flags; sf.rules = m->rules;
    sf.tail0 = std::max(0, std::min(m->rules.ts_begin, c.vocab_size)) / 32 * 32;
    sf.vpart = (float2*)m->vpart; sf.n_part = 2 * ((sf.tail0 + 127) / 128);
    sf.tail = m->logits; sf.tail_ld = (c.vocab_size - sf.tail0 + 31) / 32 * 32;
    GemmArgs g = mk(m->da, c.d_model, m->t, m->w.tok_embed, m->t, nullptr, nullptr, c.vocab_size, KW_F32, B, c.vocab_size,
                    c.d_model, EPI_ARGMAX);
    g.sample = &sf;
    static const int vocab_hint = getenv("KW_VOCAB_HINT") ? atoi(getenv("KW_VOCAB_HINT")) : 2;
    g.w_hint = vocab_hint;  // 133 MB read once per position: 2 = evict_first keeps it from displacing the layer weights
    int rc;
    {
      ProfScope ps(KW_PROF_DEC_GEMM, gemm_flops(g), st);
      rc = gemm_tc(g, st);
    }
    if (rc == KW_OK) {
      EmbedNext en;
      memset(&en, 0, sizeof(en));
      en.on = want_next && pos + 1 < c.max_target_pos && c.d_model <= 2048;
      if (en.on) {
        en.E = m->w.tok_embed; en.P = m->w.dec_pos; en.ln_w = m->dec[0].ln1_w; en.ln_b = m->dec[0].ln1_b;
        en.x = m->dx; en.da = m->da; en.d = c.d_model;
      }
      KW_TRY(sample_combine_launch(sf, tokens, B, finished, en, st));
      if (did_next) *did_next = en.on != 0;
      return KW_OK;
    }
    if (rc != KW_ERR_UNSUPPORTED) return rc;  // unsupported shape (small vocabulary): separate kernels below
  }
  KW_TRY(gemm_p(KW_PROF_DEC_GEMM, mk(m->da, c.d_model, m->t, m->w.tok_embed, m->t, nullptr, lg, c.vocab_size, KW_F32, B, c.vocab_size,
                 c.d_model, EPI_STORE), st));
  if (sample) {
    KW_REQUIRE(finished, "kw_decode_step: sample requires the finished array");
    KW_TRY(sample_launch(lg, m->flags, m->rules, tokens, ld_tokens, B, pos, begin_index, return_timestamps, finished, st));
  }
  return KW_OK;
}

int kw_decode_step(kw_model* m, int32_t* tokens, int32_t ld_tokens, int32_t B, int32_t pos, int32_t begin_index,
                   int32_t sample, int32_t return_timestamps, int32_t* finished, float* logits_out, kw_stream stream) {
  KW_TRY(bound_device_ok("kw_decode_step"));
  KW_REQUIRE(m && tokens, "kw_decode_step: null argument");
  KW_REQUIRE(B >= 1 && B <= m->cfg.max_batch && pos >= 0 && pos < m->cfg.max_target_pos && pos < ld_tokens,
             "kw_decode_step: B=%d pos=%d ld=%d out of range", B, pos, ld_tokens);
  return decode_step(m, tokens, ld_tokens, B, pos, begin_index, sample, return_timestamps, finished, logits_out,
                     (cudaStream_t)stream, false, false, nullptr);
}

int kw_sample(kw_model* m, const float* logits, int32_t* tokens, int32_t ld_tokens, int32_t B, int32_t pos,
              int32_t begin_index, int32_t return_timestamps, int32_t* finished, kw_stream stream) {
  KW_TRY(bound_device_ok("kw_sample"));
  KW_REQUIRE(m && logits && tokens && finished, "kw_sample: null argument");
  return sample_launch(logits, m->flags, m->rules, tokens, ld_tokens, B, pos, begin_index, return_timestamps, finished,
                       (cudaStream_t)stream);
}

__global__ void fill_prompt_kernel(int* tokens, int ld, int B, const int4 p0, int n_prompt, int pad, int* finished) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int pr[4] = {p0.x, p0.y, p0.z, p0.w};
  for (int j = 0; j < ld; ++j) tokens[(size_t)b * ld + j] = j < n_prompt ? pr[j] : pad;
  finished[b] = 0;
}

// `between`: called before every group of GRAPH_POS positions (kw_encode_decode slots encoder layers of another batch there)
static int greedy_pass_impl(kw_model* m, int32_t B, const int32_t* prompt, int32_t n_prompt, int32_t max_length,
                            int32_t return_timestamps, int32_t check_every, int32_t* tokens, kw_stream stream,
                            const std::function<int()>* between);

int kw_greedy_pass(kw_model* m, int32_t B, const int32_t* prompt, int32_t n_prompt, int32_t max_length,
                   int32_t return_timestamps, int32_t check_every, int32_t* tokens, kw_stream stream) {
  return greedy_pass_impl(m, B, prompt, n_prompt, max_length, return_timestamps, check_every, tokens, stream, nullptr);
}

// decoder positions per captured graph / per interleave slot (KW_GRAPH_POS: A/B measurements)
static const int GRAPH_POS = [] {
  const char* e = getenv("KW_GRAPH_POS");
  return e ? std::max(1, atoi(e)) : 8;
}();

static int greedy_pass_impl(kw_model* m, int32_t B, const int32_t* prompt, int32_t n_prompt, int32_t max_length,
                            int32_t return_timestamps, int32_t check_every, int32_t* tokens, kw_stream stream,
                            const std::function<int()>* between) {
  KW_TRY(bound_device_ok("kw_greedy_pass"));
  KW_REQUIRE(m && prompt && tokens, "kw_greedy_pass: null argument");
  KW_REQUIRE(n_prompt >= 1 && n_prompt <= 4, "kw_greedy_pass: prompt of %d tokens (1..4 supported)", n_prompt);
  KW_REQUIRE(max_length > n_prompt && max_length <= m->cfg.max_target_pos,
             "kw_greedy_pass: max_length=%d must be in (%d, %d]", max_length, n_prompt, m->cfg.max_target_pos);
  KW_REQUIRE(B >= 1 && B <= m->enc_B, "kw_greedy_pass: B=%d but encoder output holds %d rows", B, m->enc_B);
  cudaStream_t st = (cudaStream_t)stream;
  int4 p0 = make_int4(prompt[0], n_prompt > 1 ? prompt[1] : 0, n_prompt > 2 ? prompt[2] : 0, n_prompt > 3 ? prompt[3] : 0);
  fill_prompt_kernel<<<ceil_div(B, 128), 128, 0, st>>>(tokens, max_length, B, p0, n_prompt, m->rules.pad, m->finished);
  KW_LAUNCH_OK();
  ++g_launches;
  KW_TRY(kw_cross_kv(m, B, stream));
  // The all-rows-finished poll must not drain the stream (the host would then have to refill the launch queue while
  // the GPU idles, and the first kernels after the gap lose their programmatic-launch overlap): the flags are copied out
  // behind an event and looked at POLL_LAG positions later, when the copy has long completed but the GPU still has
  // those positions queued.  Positions computed after every row finished only write pad tokens over pad tokens.
  constexpr int POLL_LAG = 4;
  if (!m->finished_copied) KW_CUDA_OK(cudaEventCreateWithFlags(&m->finished_copied, cudaEventDisableTiming));
  int steps = 0, pending_since = -1;
  // one event pair around the whole position loop (no events between the decode kernels: PDL chains stay intact);
  // work = algorithmic bytes of the positions actually run: layer weights (wqkv, wo, wq_x, wo_x, w1, w2)
  // + tied vocabulary matrix + cross K/V + the self K/V rows read so far
  const double es_d = (double)esize(m->t), dd = m->cfg.d_model, Ld = m->cfg.dec_layers;
  const double w_layer_bytes = (6.0 * dd * dd + 2.0 * dd * m->cfg.ffn_dim) * es_d;
  double pass_bytes = 0.0;
  auto position_bytes = [&](int pos) {
    return Ld * w_layer_bytes + (pos >= n_prompt - 1 ? (double)m->cfg.vocab_size * dd * es_d : 0.0) +
           Ld * B * 2.0 * m->cfg.max_source_pos * dd * es_d + Ld * B * 2.0 * (pos + 1) * dd * es_d;
  };
  ProfScope pass_scope(KW_PROF_DEC_PASS, 0.0, st);
  // bf16 models, opt-in (kw_set_decode_impl(1) / KW_DECODE_FUSED=1): the whole position loop as ONE persistent kernel
  // (decode_fused.cu).  Token-identical to the schedule below up to near-ties, but measured slower on B200 (its per-phase
  // latency + grid barrier cost more than a PDL-chained launch; numbers in DESIGN.md §5), so the kernel-per-op schedule
  // stays the default.
  const bool want_fused = m->t == KW_BF16 && g_decode_impl.load() != 0 && g_gemm_impl.load() != 1;
  const int fused_rc = want_fused ? fused_decode_prepare(m) : KW_ERR_UNSUPPORTED;
  if (want_fused && fused_rc != KW_OK && (g_decode_impl.load() == 2 || getenv("KW_FUSED_REQUIRE"))) return fused_rc;
  if (fused_rc == KW_OK) {
    const int done = fused_decode_pass(m, B, n_prompt, max_length, return_timestamps, tokens, st);
    if (done < 0) return done;
    for (int pos = 0; pos < done; ++pos) pass_bytes += position_bytes(pos);
    if (g_prof.mask & (1u << KW_PROF_DEC_PASS)) g_prof.work[KW_PROF_DEC_PASS] += pass_bytes;
    return done;
  }
  // CUDA-graph replay: the kernel-per-op schedule of GRAPH_POS consecutive positions is captured once (on a private
  // stream — the caller's may be the legacy default stream, which cannot capture) and re-launched on every later pass
  // with the same shape; programmatic-launch edges are kept by the capture.  Not used while per-kernel profiling events
  // or the in-kernel timeline stamps are active (they would be baked into the graph).
  const unsigned per_kernel_prof = (1u << KW_PROF_DEC_GEMM) | (1u << KW_PROF_DEC_CROSS);
  if (g_decode_graph.load() && !(g_prof.mask & per_kernel_prof) && !getenv("KW_DECODE_SKIP")) {
    const int cfg_epoch = g_sample_fused.load() | (g_gemm_impl.load() << 1);
    if (!m->cap_stream) KW_CUDA_OK(cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking));
    int* gt = m->gtokens;
    fill_prompt_kernel<<<ceil_div(B, 128), 128, 0, st>>>(gt, max_length, B, p0, n_prompt, m->rules.pad, m->finished);
    KW_LAUNCH_OK();
    ++g_launches;
    bool pre = false, stop = false;
    int pos = 0;
    while (pos + 1 < max_length && !stop) {
      if (between) KW_TRY((*between)());
      const int end = std::min(pos + GRAPH_POS, max_length - 1);
      kw_model::PassGraph* pg = nullptr;
      for (auto& g : m->graphs)
        if (g.B == B && g.n_prompt == n_prompt && g.max_length == max_length && g.return_ts == return_timestamps &&
            g.pos0 == pos && g.pos1 == end && g.pre_in == (int)pre && g.cfg_epoch == cfg_epoch) {
          pg = &g;
          break;
        }
      if (!pg) {
        if (m->graphs.size() >= 512) {  // shapes keep changing (ragged last batches): start over instead of growing
          for (auto& g : m->graphs) cudaGraphExecDestroy(g.exec);
          m->graphs.clear();
        }
        kw_model::PassGraph ng = {B, n_prompt, max_length, return_timestamps, pos, end, (int)pre, cfg_epoch, nullptr, 0, false};
        const long long l0 = g_launches.load();
        KW_CUDA_OK(cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal));
        int rc = KW_OK;
        bool pr = pre;
        for (int q = pos; q < end && rc == KW_OK; ++q) {
          bool did = false;
          rc = decode_step(m, gt, max_length, B, q, n_prompt, q >= n_prompt - 1, return_timestamps, m->finished, nullptr,
                           m->cap_stream, pr, q + 2 < max_length, &did);
          pr = did;
        }
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamEndCapture(m->cap_stream, &graph);
        ng.n_launch = (int)(g_launches.load() - l0);
        g_launches.store(l0);
        if (rc != KW_OK) {
          if (graph) cudaGraphDestroy(graph);
          return rc;
        }
        if (ce != cudaSuccess || !graph) {
          set_error("kw_greedy_pass: graph capture failed: %s", cudaGetErrorString(ce));
          return KW_ERR_CUDA;
        }
        ce = cudaGraphInstantiate(&ng.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) {
          set_error("kw_greedy_pass: graph instantiation failed: %s", cudaGetErrorString(ce));
          return KW_ERR_CUDA;
        }
        ng.pre_out = pr;
        m->graphs.push_back(ng);
        pg = &m->graphs.back();
      }
      KW_CUDA_OK(cudaGraphLaunch(pg->exec, st));
      g_launches += pg->n_launch;
      pre = pg->pre_out;
      for (int q = pos; q < end; ++q) pass_bytes += position_bytes(q);
      steps += end - pos;
      pos = end;
      // all-rows-finished poll with one graph of lag: the flags copied out behind graph i are looked at after graph
      // i + 1 has been enqueued, so the GPU never waits for the host
      if (check_every > 0 && pos + 1 < max_length) {
        if (pending_since >= 0) {
          KW_CUDA_OK(cudaEventSynchronize(m->finished_copied));
          bool all = true;
          for (int b = 0; b < B; ++b) all = all && m->finished_host[b];
          stop = all;
        }
        if (!stop && pos >= n_prompt) {
          KW_CUDA_OK(cudaMemcpyAsync(m->finished_host, m->finished, B * sizeof(int), cudaMemcpyDeviceToHost, st));
          KW_CUDA_OK(cudaEventRecord(m->finished_copied, st));
          pending_since = pos;
        }
      }
    }
    if (pending_since >= 0) KW_CUDA_OK(cudaEventSynchronize(m->finished_copied));
    KW_CUDA_OK(cudaMemcpyAsync(tokens, gt, (size_t)B * max_length * sizeof(int), cudaMemcpyDeviceToDevice, st));
    if (g_prof.mask & (1u << KW_PROF_DEC_PASS)) g_prof.work[KW_PROF_DEC_PASS] += pass_bytes;
    return steps;
  }
  bool pre_embedded = false;
  for (int pos = 0; pos + 1 < max_length; ++pos) {
    if (between && pos % GRAPH_POS == 0) KW_TRY((*between)());
    const int sample = pos >= n_prompt - 1;
    bool did_next = false;
    KW_TRY(decode_step(m, tokens, max_length, B, pos, n_prompt, sample, return_timestamps, m->finished, nullptr, st,
                       pre_embedded, pos + 2 < max_length, &did_next));
    pre_embedded = did_next;
    ++steps;
    pass_bytes += position_bytes(pos);
    const int generated = pos + 2 - n_prompt;  // tokens sampled so far
    if (check_every > 0 && sample && generated % check_every == 0 && pos + 2 < max_length && pending_since < 0) {
      KW_CUDA_OK(cudaMemcpyAsync(m->finished_host, m->finished, B * sizeof(int), cudaMemcpyDeviceToHost, st));
      KW_CUDA_OK(cudaEventRecord(m->finished_copied, st));
      pending_since = pos;
    }
    if (pending_since >= 0 && pos >= pending_since + POLL_LAG) {
      KW_CUDA_OK(cudaEventSynchronize(m->finished_copied));
      pending_since = -1;
      bool all = true;
      for (int b = 0; b < B; ++b) all = all && m->finished_host[b];
      if (all) break;
    }
  }
  if (pending_since >= 0) KW_CUDA_OK(cudaEventSynchronize(m->finished_copied));  // finished_host is reused by the next pass
  if (g_prof.mask & (1u << KW_PROF_DEC_PASS)) g_prof.work[KW_PROF_DEC_PASS] += pass_bytes;
  return steps;
}

// Two batches in flight on one stream: the encoder of batch i + 1 (model `enc`) in layer groups, with the decoder
// positions of batch i (model `dec`, whose encoder output is already in place) slotted between the groups, GRAPH_POS
// positions at a time.  Same kernels and the same results as kw_encode(enc) + kw_greedy_pass(dec); what changes is the
// power profile the clock governor sees: a ~140 ms block of power-capped GEMMs followed by a ~45 ms block of latency-bound
// decode kernels leaves the SM clock low for the whole decode block, alternating ~9 ms / ~3 ms slices does not
// (tools/interleave_probe.py).  `enc` and `dec` are two kw_model handles created over the same weight table (two
// workspaces); dec == NULL runs the encoder alone (first batch of a stream).
int kw_encode_decode(kw_model* enc, const float* mel, int32_t B_enc, kw_model* dec, int32_t B_dec, const int32_t* prompt,
                     int32_t n_prompt, int32_t max_length, int32_t return_timestamps, int32_t check_every,
                     int32_t* tokens, kw_stream stream) {
  KW_TRY(bound_device_ok("kw_encode_decode"));
  KW_REQUIRE(enc && mel, "kw_encode_decode: null argument");
  KW_REQUIRE(enc != dec, "kw_encode_decode: the two batches need two model handles (two workspaces)");
  KW_REQUIRE(B_enc >= 1 && B_enc <= enc->cfg.max_batch, "kw_encode_decode: B_enc=%d outside [1, max_batch=%d]", B_enc,
             enc->cfg.max_batch);
  cudaStream_t st = (cudaStream_t)stream;
  if (!dec) return kw_encode(enc, mel, B_enc, nullptr, stream);
  KW_REQUIRE(prompt && tokens, "kw_encode_decode: null argument");
  const int L = enc->cfg.enc_layers;
  const int slots = std::max(1, (max_length - 1 + GRAPH_POS - 1) / GRAPH_POS);
  const int per_slot = (L + slots - 1) / slots;
  int next_layer = 0;
  bool stem_done = false;
  const std::function<int()> between = [&]() -> int {
    if (!stem_done) {
      KW_TRY(encode_stem(enc, mel, B_enc, st));
      stem_done = true;
    }
    const int l1 = std::min(L, next_layer + per_slot);
    KW_TRY(encode_layers(enc, B_enc, next_layer, l1, st));
    next_layer = l1;
    return KW_OK;
  };
  const int steps = greedy_pass_impl(dec, B_dec, prompt, n_prompt, max_length, return_timestamps, check_every, tokens,
                                     stream, &between);
  if (steps < 0) return steps;
  if (!stem_done) KW_TRY(encode_stem(enc, mel, B_enc, st));
  KW_TRY(encode_layers(enc, B_enc, next_layer, L, st));  // the decoder stopped early (every row finished)
  KW_TRY(encode_finish(enc, B_enc, nullptr, st));
  return steps;
}

// Teacher-forcing decoder forward over T positions at once (no cache): WhisperDecoder.forward + proj_out with
// decoder_input_ids [B, T] (modeling_whisper.py:734-796, 449-506, 1069-1081), as run_distillation.py:641-649 calls the
// frozen teacher.  Rows are (b, t)-major; the encoder workspaces double as decoder workspaces (B*T <= B*1500).
int kw_decoder_forward(kw_model* m, const int32_t* decoder_input_ids, int32_t B, int32_t T, float* logits_out,
                       kw_stream stream) {
  KW_TRY(bound_device_ok("kw_decoder_forward"));
  KW_REQUIRE(m && decoder_input_ids && logits_out, "kw_decoder_forward: null argument");
  KW_REQUIRE(B >= 1 && B <= m->enc_B, "kw_decoder_forward: B=%d but encoder output holds %d rows", B, m->enc_B);
  KW_REQUIRE(T >= 1 && T <= m->cfg.max_target_pos, "kw_decoder_forward: T=%d outside [1, %d]", T, m->cfg.max_target_pos);
  cudaStream_t st = (cudaStream_t)stream;
  const kw_config& c = m->cfg;
  const kw_dtype t = m->t;
  const int d = c.d_model, S = c.max_source_pos, F = c.ffn_dim, H = c.n_heads, M = B * T;
  const size_t es = esize(t);
  const size_t xkv_stride = (size_t)c.max_batch * S * 2 * d * es;
  KW_TRY(kw_cross_kv(m, B, stream));
  KW_TRY(embed_seq(decoder_input_ids, T, T, m->w.tok_embed, m->w.dec_pos, m->x, B, d, c.vocab_size, t, st));
  for (int l = 0; l < c.dec_layers; ++l) {
    const kw_dec_layer_weights& w = m->dec[l];
    KW_TRY(layernorm(m->x, w.ln1_w, w.ln1_b, m->a, M, d, t, st));
    KW_TRY(gemm(mk(m->a, d, t, w.wqkv, t, w.bqkv, m->bufP, 3 * d, t, M, 3 * d, d, EPI_STORE), st));
    const char* qkv = (const char*)m->bufP;
    KW_TRY(attention_simt(qkv, qkv + d * es, qkv + 2 * d * es, m->o, B, H, T, T, (long long)T * 3 * d, 3 * d,
                          (long long)T * 3 * d, 3 * d, (long long)T * d, d, t, st, /*causal=*/1));
    KW_TRY(gemm(mk(m->o, d, t, w.wo, t, w.bo, m->x, d, KW_F32, M, d, d, EPI_RESID), st));
    KW_TRY(layernorm(m->x, w.lnx_w, w.lnx_b, m->a, M, d, t, st));
    KW_TRY(gemm(mk(m->a, d, t, w.wq_x, t, w.bq_x, m->bufP, d, t, M, d, d, EPI_STORE), st));
    const char* xkv = (const char*)m->xkv + l * xkv_stride;
    KW_TRY(kw_attention(m->bufP, xkv, xkv + d * es, m->o, B, H, T, S, (long long)T * d, d, (long long)S * 2 * d, 2 * d,
                        (long long)T * d, d, t, stream));
    KW_TRY(gemm(mk(m->o, d, t, w.wo_x, t, w.bo_x, m->x, d, KW_F32, M, d, d, EPI_RESID), st));
    KW_TRY(layernorm(m->x, w.ln3_w, w.ln3_b, m->a, M, d, t, st));
    KW_TRY(gemm(mk(m->a, d, t, w.w1, t, w.b1, m->bufQ, F, t, M, F, d, EPI_GELU), st));
    KW_TRY(gemm(mk(m->bufQ, F, t, w.w2, t, w.b2, m->x, d, KW_F32, M, d, F, EPI_RESID), st));
  }
  KW_TRY(layernorm(m->x, m->w.dec_ln_w, m->w.dec_ln_b, m->a, M, d, t, st));
  // proj_out over all positions.  The vocabulary (51866) is not a multiple of the wide tcgen05 kernel's 32-column
  // granule, so the bf16 path streams it through the decode-time (weights-stationary) kernel 64 rows at a time.
  if (t == KW_BF16 && c.vocab_size % 32 != 0 && g_gemm_impl.load() != 1) {
    for (int r0 = 0; r0 < M; r0 += 64) {
      const int rows = std::min(64, M - r0);
      KW_TRY(gemm(mk((const char*)m->a + (size_t)r0 * d * es, d, t, m->w.tok_embed, t, nullptr,
                     logits_out + (size_t)r0 * c.vocab_size, c.vocab_size, KW_F32, rows, c.vocab_size, d, EPI_STORE), st));
    }
  } else {
    KW_TRY(gemm(mk(m->a, d, t, m->w.tok_embed, t, nullptr, logits_out, c.vocab_size, KW_F32, M, c.vocab_size, d,
                   EPI_STORE), st));
  }
  return KW_OK;
}

int kw_attention(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t H, int32_t Tq, int32_t Tk,
                 int64_t q_stride_b, int64_t q_stride_t, int64_t kv_stride_b, int64_t kv_stride_t, int64_t o_stride_b,
                 int64_t o_stride_t, int32_t dtype, kw_stream stream) {
  KW_TRY(bound_device_ok("kw_attention"));
  KW_REQUIRE(q && k && v && out, "kw_attention: null pointer");
  if (dtype == KW_BF16 && g_gemm_impl.load() != 1) {
    int rc = attention_tc(q, k, v, out, B, H, Tq, Tk, q_stride_b, q_stride_t, kv_stride_b, kv_stride_t, o_stride_b,
                          o_stride_t, (cudaStream_t)stream);
    if (rc != KW_ERR_UNSUPPORTED) return rc;
  }
  return attention_simt(q, k, v, out, B, H, Tq, Tk, q_stride_b, q_stride_t, kv_stride_b, kv_stride_t, o_stride_b,
                        o_stride_t, (kw_dtype)dtype, (cudaStream_t)stream);
}

int kw_linear(const void* A, const void* W, const float* bias, void* out, int32_t M, int32_t N, int32_t K, int32_t epi,
              int32_t a_dtype, int32_t w_dtype, int32_t out_dtype, int32_t impl, kw_stream stream) {
  KW_TRY(bound_device_ok("kw_linear"));
  KW_REQUIRE(A && W && out && epi >= 0 && epi <= 2, "kw_linear: bad arguments");
  GemmArgs g = mk(A, K, (kw_dtype)a_dtype, W, (kw_dtype)w_dtype, bias, out, N, (kw_dtype)out_dtype, M, N, K, epi);
  if (impl == 1) return gemm_simt(g, (cudaStream_t)stream);
  if (impl == 2) return gemm_tc(g, (cudaStream_t)stream);
  return gemm(g, (cudaStream_t)stream);
}

int kw_layernorm(const float* x, const float* w, const float* b, void* out, int32_t rows, int32_t d, int32_t out_dtype,
                 kw_stream stream) {
  KW_TRY(bound_device_ok("kw_layernorm"));
  KW_REQUIRE(x && w && b && out, "kw_layernorm: null pointer");
  return layernorm(x, w, b, out, rows, d, (kw_dtype)out_dtype, (cudaStream_t)stream);
}

}  // extern "C"
