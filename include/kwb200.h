/* kwb200 — C ABI of the B200-native Whisper transcription hot path (libkwb200.so).
 *
 * The reference (kotoba-tech/kotoba-whisper) has no native code and no FFI: its hot path is three Python calls into
 * `transformers`.  Each entry point below names the Python interface it stands behind; the Python drop-in
 * (kotoba_whisper_b200/) binds these with ctypes and re-exposes the reference's call surface unchanged.
 *
 *   kw_logmel            <- WhisperFeatureExtractor.__call__ / _torch_extract_fbank_features
 *                           (HF/models/whisper/feature_extraction_whisper.py:189-342, 135-164; called from
 *                            run_pseudo_labelling.py:268, run_data_filtering.py:338, HF pipelines/automatic_speech_recognition.py:67-72)
 *   kw_encode            <- WhisperEncoder.forward (HF/models/whisper/modeling_whisper.py:593-647), reached from
 *                           generate() via GenerationMixin._prepare_encoder_decoder_kwargs_for_generation (generation/utils.py:765-804)
 *   kw_cross_kv          <- cross-attention K/V build inside WhisperAttention.forward (modeling_whisper.py:326-336)
 *   kw_decode_step       <- WhisperDecoder.forward with EncoderDecoderCache, one token (modeling_whisper.py:734-796, 449-506)
 *                           + proj_out (:1081) + the three Whisper logits processors (generation/logits_process.py:1855-2043)
 *                           + argmax / eos / pad bookkeeping of GenerationMixin._sample (generation/utils.py:2762-2805)
 *   kw_greedy_pass       <- one GenerationMixin.generate call as issued by generate_with_fallback
 *                           (HF/models/whisper/generation_whisper.py:1027): encoder output -> prompt prefill -> greedy loop
 *   kw_attention         <- the attention plug-in seam, ALL_ATTENTION_FUNCTIONS[name](module, q, k, v, ...) as called at
 *                           modeling_whisper.py:342-352 (what `attn_implementation="sdpa"|"flash_attention_2"` selects,
 *                           run_pseudo_labelling.py:64,230)
 *
 * Conventions: every pointer marked `dev` is a device pointer owned by the caller (torch tensors in the Python host);
 * nothing is allocated per call (workspaces and KV pools are created in kw_model_create for `max_batch`); every call
 * enqueues work on the given stream and does NOT synchronise unless stated; functions return 0 on success or a
 * negative kw_status, with a message available from kw_last_error(); no exceptions cross the ABI; a model handle is
 * not thread-safe (one host thread per GPU process, as under `accelerate launch`).
 */
#ifndef KWB200_H
#define KWB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* kw_stream; /* cudaStream_t */

typedef enum { KW_OK = 0, KW_ERR_ARG = -1, KW_ERR_CUDA = -2, KW_ERR_UNSUPPORTED = -3, KW_ERR_NOMEM = -4 } kw_status;
typedef enum { KW_F32 = 0, KW_BF16 = 1 } kw_dtype;

/* Model dimensions = the WhisperConfig fields the path reads (HF/models/whisper/configuration_whisper.py). */
typedef struct {
  int32_t vocab_size;      /* 51866 */
  int32_t n_mels;          /* 128 (80 for v1/v2-era checkpoints) */
  int32_t d_model;         /* 1280; multiple of 64 */
  int32_t n_heads;         /* 20;  d_model / n_heads must be 64 */
  int32_t ffn_dim;         /* 5120 */
  int32_t enc_layers;      /* 32 */
  int32_t dec_layers;      /* 2 (kotoba / distil) or 32 (teacher) */
  int32_t max_source_pos;  /* 1500 */
  int32_t max_target_pos;  /* 448 */
  int32_t dtype;           /* kw_dtype of weight matrices, activations and KV caches; accumulation is always fp32 */
  int32_t max_batch;       /* workspaces are sized for this many 30 s windows */
} kw_config;

/* Weight table: device pointers into caller-owned tensors, already in kernel layout (the Python repacker builds it
 * from an HF state_dict, names in SURVEY.md §8b).  Matrices are `dtype` ([out, in] row-major like nn.Linear);
 * biases, LayerNorm parameters and position tables are always fp32. */
typedef struct {
  const float *ln1_w, *ln1_b; /* self_attn_layer_norm */
  const void* wqkv;           /* [3d, d] = [q*0.125 ; k ; v]  (the 1/sqrt(64) q scale is folded in: exact, power of 2) */
  const float* bqkv;          /* [3d]    = [bq*0.125 ; 0 ; bv] (k_proj has no bias, modeling_whisper.py:279) */
  const void* wo;             /* [d, d] */
  const float* bo;
  const float *ln2_w, *ln2_b; /* final_layer_norm */
  const void* w1;             /* [ffn, d] */
  const float* b1;
  const void* w2;             /* [d, ffn] */
  const float* b2;
} kw_enc_layer_weights;

typedef struct {
  const float *ln1_w, *ln1_b; /* self_attn_layer_norm */
  const void* wqkv;           /* [3d, d], q pre-scaled as above */
  const float* bqkv;
  const void* wo;
  const float* bo;
  const float *lnx_w, *lnx_b; /* encoder_attn_layer_norm */
  const void* wq_x;           /* [d, d] cross-attention q (pre-scaled) */
  const float* bq_x;
  const void* wkv_x;          /* [2d, d] = [k ; v] applied to the encoder output once per pass */
  const float* bkv_x;         /* [2d]    = [0 ; bv] */
  const void* wo_x;
  const float* bo_x;
  const float *ln3_w, *ln3_b; /* final_layer_norm */
  const void* w1;
  const float* b1;
  const void* w2;
  const float* b2;
} kw_dec_layer_weights;

typedef struct {
  const void* conv1_w;  /* [d, 3*n_mels], column = tap*n_mels + mel   (from conv1.weight [d, n_mels, 3]) */
  const float* conv1_b;
  const void* conv2_w;  /* [d, 3*d],      column = tap*d + channel    (from conv2.weight [d, d, 3]) */
  const float* conv2_b;
  const float* enc_pos; /* [max_source_pos, d] sinusoid table (embed_positions.weight) */
  const float *enc_ln_w, *enc_ln_b;
  const void* tok_embed; /* [vocab, d]; also proj_out (tied, modeling_whisper.py:966) */
  const float* dec_pos;  /* [max_target_pos, d] learned */
  const float *dec_ln_w, *dec_ln_b;
  const kw_enc_layer_weights* enc; /* host array [enc_layers] */
  const kw_dec_layer_weights* dec; /* host array [dec_layers] */
} kw_weights;

/* Token rules = the GenerationConfig fields read by the three processors and by _sample. */
typedef struct {
  int32_t eos_token_id;                /* 50257 */
  int32_t pad_token_id;                /* 50257 */
  int32_t no_timestamps_token_id;      /* 50364; timestamp_begin = this + 1 */
  int32_t max_initial_timestamp_index; /* 50; < 0 = None */
  const int32_t* suppress_tokens;      /* host array */
  int32_t n_suppress;
  const int32_t* begin_suppress_tokens; /* host array */
  int32_t n_begin_suppress;
} kw_token_rules;

typedef struct kw_model kw_model;

const char* kw_last_error(void);
const char* kw_version(void);

/* ---- log-mel ---------------------------------------------------------------------------------------------------
 * audio  dev f32 [B, n_samples]   clip b occupies audio[b*n_samples ...]; samples >= lens[b] are read as 0 (right pad)
 * lens   dev i32 [B] or NULL      (NULL: every clip is n_samples long)
 * out    dev f32 [B, n_mels, n_samples/160]   (mel-major, time contiguous — HF's `input_features` layout;
 *                                               n_samples/160 rounds down: HF computes 1 + n/160 frames and drops the last)
 * n_samples >= 400; n_mels <= 128 with n_mels * (n_samples/160) a multiple of 4 (80 and 128 always qualify).
 * clip_max dev f32 [B] scratch (per-clip max of log10 mel), overwritten. */
int kw_logmel(const float* audio, const int32_t* lens, int32_t B, int32_t n_samples, int32_t n_mels, float* out,
              float* clip_max, kw_stream stream);
/* Window mode — the device-side chunker of the ASR pipeline (HF/pipelines/automatic_speech_recognition.py:61-84
 * `chunk_iter`, called on behalf of run_speed_eval.py:76): the recording is uploaded ONCE; window w covers samples
 * [starts[w], starts[w] + lens[w]) of it and is featurised as if it had been copied out and right-padded with zeros
 * to n_samples (lens[w] <= n_samples; the caller guarantees starts[w] + lens[w] <= recording length).
 * recording dev f32 [n_total]; starts dev i64 [W]; lens dev i32 [W]; out dev f32 [W, n_mels, n_samples/160]. */
int kw_logmel_windows(const float* recording, const int64_t* starts, const int32_t* lens, int32_t W, int32_t n_samples,
                      int32_t n_mels, float* out, float* clip_max, kw_stream stream);
/* Slaney filterbank as the kernel uses it, float64 [201, n_mels] row-major, written to a HOST buffer (test hook). */
int kw_mel_filterbank(int32_t n_mels, double* out_host);

/* ---- model ------------------------------------------------------------------------------------------------------ */
int kw_model_create(const kw_config* cfg, const kw_weights* w, const kw_token_rules* rules, kw_model** out);
void kw_model_destroy(kw_model* m);
/* bytes of device memory the handle allocated (workspaces + KV pools) */
int64_t kw_model_workspace_bytes(const kw_model* m);

/* mel dev f32 [B, n_mels, 2*max_source_pos] -> encoder output kept inside the handle; if enc_out (dev f32
 * [B, max_source_pos, d]) is non-NULL a copy of the final-LayerNorm output is written there. */
int kw_encode(kw_model* m, const float* mel, int32_t B, float* enc_out, kw_stream stream);
/* Use a caller-provided encoder output (generate(encoder_outputs=...)): dev f32 [B, max_source_pos, d]. */
int kw_set_encoder_output(kw_model* m, const float* enc, int32_t B, kw_stream stream);
/* Project the handle's encoder output to every decoder layer's cross-attention K/V pool. */
int kw_cross_kv(kw_model* m, int32_t B, kw_stream stream);

/* One decoder position for B rows.
 * tokens dev i32 [B, ld_tokens]  token history: columns [0, pos] are read (column pos = the token fed in this step);
 *                                when sample != 0 the chosen next token is written to column pos+1
 * pos                            index of the fed token in the decoder sequence (0 = <|startoftranscript|>)
 * begin_index                    prompt length (first generated token is at column begin_index)
 * sample                         0: only extend the KV cache (prompt prefill); 1: also logits -> processors -> argmax
 * return_timestamps              enables the WhisperTimeStamp rules
 * finished dev i32 [B]           row state (1 after eos was emitted; such rows emit pad); updated when sample != 0
 * logits_out dev f32 [B, vocab] or NULL   raw (unprocessed) fp32 logits of this step, for parity tests */
int kw_decode_step(kw_model* m, int32_t* tokens, int32_t ld_tokens, int32_t B, int32_t pos, int32_t begin_index,
                   int32_t sample, int32_t return_timestamps, int32_t* finished, float* logits_out, kw_stream stream);

/* Whole greedy pass on the handle's current encoder output (kw_encode / kw_set_encoder_output must have run):
 * cross K/V, prefill of `n_prompt` prompt tokens (host array), then greedy steps until every row has emitted eos or
 * the sequence length reaches max_length.  tokens dev i32 [B, max_length] receives prompt + generated ids (pad after
 * eos).  Synchronises the stream every `check_every` steps to test the all-finished flag (0 = never, run to
 * max_length).  Returns the number of decoder positions evaluated (>= 0) or a negative kw_status.
 * B <= max_batch.  The bf16 decode-time kernels take up to 128 rows per launch (the batch is the N of the weight-streaming
 * tcgen05 GEMMs: 64- and 128-wide instantiations); a position costs nearly the same for 128 rows as for 64 outside the
 * cross-attention K/V stream, which is why the host side coalesces two 64-utterance batches into one 128-row pass
 * (GenerateStream(coalesce=2)).  A row's ids do not depend on the batch it is decoded in. */
int kw_greedy_pass(kw_model* m, int32_t B, const int32_t* prompt, int32_t n_prompt, int32_t max_length,
                   int32_t return_timestamps, int32_t check_every, int32_t* tokens, kw_stream stream);

/* Two batches in flight (throughput mode of the labelling loop, run_pseudo_labelling.py:333-341: `for batch in loader:
 * generate(batch)`): the encoder of the NEXT batch (`enc`, features `mel` dev f32 [B_enc, n_mels, 2*max_source_pos]) runs
 * in layer groups with the greedy pass of the PREVIOUS batch (`dec`, encoder output already in place; arguments as
 * kw_greedy_pass) slotted between the groups on the same stream.  Same kernels and results as kw_encode(enc) followed
 * by kw_greedy_pass(dec).  `enc` and `dec` must be two handles (kw_model_create twice over the same weight table: two
 * workspaces, one copy of the weights).  dec == NULL: encoder only.  Returns the number of decoder positions evaluated
 * (0 when dec == NULL) or a negative kw_status. */
int kw_encode_decode(kw_model* enc, const float* mel, int32_t B_enc, kw_model* dec, int32_t B_dec, const int32_t* prompt,
                     int32_t n_prompt, int32_t max_length, int32_t return_timestamps, int32_t check_every,
                     int32_t* tokens, kw_stream stream);

/* Teacher-forcing decoder forward (no KV cache) on the handle's current encoder output: all T positions at once,
 * causal self-attention, cross-attention over the encoder K/V, proj_out — what `teacher_model(encoder_outputs=...,
 * labels=...)` computes in the reference's distillation step (run_distillation.py:641-649; WhisperDecoder.forward,
 * HF/models/whisper/modeling_whisper.py:734-796, 1069-1081).
 * decoder_input_ids dev i32 [B, T] (already shifted right by the caller); logits_out dev f32 [B, T, vocab]. */
int kw_decoder_forward(kw_model* m, const int32_t* decoder_input_ids, int32_t B, int32_t T, float* logits_out,
                       kw_stream stream);

/* ---- attention seam ------------------------------------------------------------------------------------------------
 * q,k,v dev [B, T, 3?]: strided views, element (b, t, h, e) at  ptr + b*stride_b + t*stride_t + h*64 + e ; head dim 64,
 * softmax(q k^T) v with q pre-scaled (scaling = 1.0 at modeling_whisper.py:349), no mask.  out dev [B, Tq, H*64]. */
int kw_attention(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t H, int32_t Tq, int32_t Tk,
                 int64_t q_stride_b, int64_t q_stride_t, int64_t kv_stride_b, int64_t kv_stride_t, int64_t o_stride_b,
                 int64_t o_stride_t, int32_t dtype, kw_stream stream);

/* ---- building blocks exported for the parity tests -------------------------------------------------------------- */
/* out[M,N] = epilogue(A[M,K] . W[N,K]^T + bias); epi: 0 store, 1 gelu, 2 out(f32) += ; impl: 0 auto, 1 simt, 2 tcgen05 */
int kw_linear(const void* A, const void* W, const float* bias, void* out, int32_t M, int32_t N, int32_t K, int32_t epi,
              int32_t a_dtype, int32_t w_dtype, int32_t out_dtype, int32_t impl, kw_stream stream);
int kw_layernorm(const float* x, const float* w, const float* b, void* out, int32_t rows, int32_t d, int32_t out_dtype,
                 kw_stream stream);
/* processors + argmax on given fp32 logits (same kernel kw_decode_step uses) */
int kw_sample(kw_model* m, const float* logits, int32_t* tokens, int32_t ld_tokens, int32_t B, int32_t pos,
              int32_t begin_index, int32_t return_timestamps, int32_t* finished, kw_stream stream);

/* ---- optional device timing of kernel categories (CUDA events on the launching stream) ---------------------------
 * kw_profile_enable(mask) turns on event pairs around every launch of the categories in `mask` (bit = category);
 * kw_profile_read returns the summed device time, launch count and ALGORITHMIC work (FLOPs for GEMM / attention
 * categories, bytes for the memory-bound ones) since the last reset; it synchronises on the recorded events. */
enum {
  KW_PROF_ENC_GEMM = 0,  /* conv-stem, QKV, out, fc1, fc2 projections of kw_encode (FLOPs) */
  KW_PROF_ENC_ATTN = 1,  /* encoder self-attention (FLOPs: 4 B H S S 64) */
  KW_PROF_XKV_GEMM = 2,  /* one-shot cross K/V projection (FLOPs) */
  KW_PROF_DEC_GEMM = 3,  /* decode-step projections incl. vocab (FLOPs) */
  KW_PROF_DEC_CROSS = 4, /* decode-step cross-attention (bytes: cached K and V read once) */
  KW_PROF_LOGMEL = 5,    /* log-mel (bytes: audio in + features out) */
  KW_PROF_DEC_PASS = 6,  /* all decoder positions of one kw_greedy_pass under ONE event pair (bytes: per position
                            actually run, the layer weights + vocabulary matrix + cross K/V + self K/V rows read so
                            far) */
  KW_PROF_NCAT = 7
};
void kw_profile_enable(uint32_t category_mask);
int kw_profile_read(int32_t category, double* total_ms, int64_t* launches, double* work, int32_t reset);

/* Bring-up hook: shared-memory descriptor parameters (bytes) of the MN-major V operand in the tcgen05 attention kernel
 * (defaults 16 / 1024 / 2048).  Not part of the stable surface. */
void kw_debug_attention_desc(int32_t v_lbo_bytes, int32_t v_sbo_bytes, int32_t v_kstep_bytes);

/* Bring-up hook: when non-NULL, CTA 0 of the decode-time GEMM writes %globaltimer stamps (ns) of its pipeline phases
 * into this device buffer of 16 uint64 (0 start, 1 weights requested, 2 dependency wait done, 3 first stage landed,
 * 4 last MMA committed, 8 accumulator visible to the epilogue, 5 epilogue done, 7 all warps done). */
void kw_debug_gemm_stamps(uint64_t* dev_buffer_16);

/* 0: auto (tcgen05 where eligible), 1: force SIMT GEMMs, 2: force tcgen05 (error if ineligible). Process-wide. */
void kw_set_gemm_impl(int32_t impl);
/* 1: wide GEMMs (M >= 256) use the 2-CTA tcgen05 kernel (cta_group::2, 256 x 256 tiles per CTA pair); 0: 1-CTA kernel. */
void kw_set_gemm_2cta(int32_t on);
/* Decode schedule of kw_greedy_pass.  0 (default): one kernel per op, chained with programmatic dependent launch.
 * 1: a bf16 model runs the whole position loop as ONE persistent cooperative kernel (csrc/decode_fused.cu: LayerNorm +
 * QKV, self-attention with KV append, cross-attention, GELU MLP, vocabulary projection with the logits processors and
 * argmax in its epilogue, grid barriers between phases) when its shape fits, else falls back; 2: as 1 but an error
 * instead of the fallback.  Same tokens up to near-ties; opt-in because it measures slower on B200.
 * Process-wide; initial value from the environment variable KW_DECODE_FUSED. */
void kw_set_decode_impl(int32_t impl);
/* 1 (default): in kw_decode_step with sample != 0 and no logits_out, a bf16 model's vocabulary projection applies the
 * three logits processors and the arg-max in its own epilogue (per-warp partials + a one-CTA-per-row combine kernel):
 * the [B, vocab] fp32 logits are never written or re-read.  0: projection -> fp32 logits -> sample kernel.  Process-wide;
 * initial value from KW_SAMPLE_FUSED. */
void kw_set_sample_fused(int32_t on);
/* 1 (default): kw_greedy_pass replays the decoder positions as CUDA graphs — the kernel-per-op schedule of 8 consecutive
 * positions captured once per (batch, prompt length, max_length, timestamps) and re-launched on later passes, tokens kept
 * in a library-owned buffer and copied to the caller's at the end; the all-rows-finished poll runs between graphs.
 * 0: every kernel launched from the host each pass.  Same kernels, same tokens.  Process-wide; initial value from
 * KW_DECODE_GRAPH. */
void kw_set_decode_graph(int32_t on);
/* counts kernels launched by this library since the last reset (bench.py's gpu_launches) */
int64_t kw_launch_count(int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* KWB200_H */
