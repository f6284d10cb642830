// Fused STFT -> power -> mel -> log10 kernel (+ per-clip clamp/affine fix-up) for Whisper log-mel features.
//
// Arithmetic contract (HF/models/whisper/feature_extraction_whisper.py:135-164, filterbank HF/audio_utils.py:453-544):
//   frames of 400 samples, hop 160, centre reflect padding, periodic Hann, 201-bin power spectrum, frame n/160 dropped,
//   Slaney mel projection, log10(max(.,1e-10)), max(., clipmax-8), (.+4)/4  with clipmax per clip.
//
// Layout / algorithm: one CTA owns FT consecutive frames of one clip.  The (160*FT + 240) samples they touch are
// staged once in shared memory (each sample is used by 2.5 frames), then groups of 40 threads run one 400-point real
// FFT each as a 200-point complex Stockham FFT (radices 5,5,8) in float64 (the 1e-5 abs parity bar is not reachable
// with an fp32 transform: single-bin mel filters amplify the rounding of near-empty bins through log10),
// fold it to 201 real-input bins, apply the sparse (<= 16 taps) filterbank and write a [n_mels][FT] tile so global
// stores are contiguous along time.  clipmax is reduced with one atomic per CTA; a second, L2-resident pass applies
// the clamp and the affine map in place.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace kw {

constexpr int NFFT = 400, HOP = 160, NBIN = 201, NC = 200;
constexpr int FT = 40;          // frames per CTA (3000 = 75 * 40)
constexpr int TPF = 40;         // threads cooperating on one frame
constexpr int FPI = 8;          // frames in flight per iteration
constexpr int NTHREADS = TPF * FPI;
constexpr int NSAMP = HOP * (FT - 1) + NFFT;  // samples staged per CTA
constexpr int MAX_TAPS = 16, MAX_MELS = 128;

constexpr int MAX_TOTAL_TAPS = 512;  // non-zeros of the whole filterbank (every bin feeds <= 2 filters: ~2 * 201)
struct MelBankDev {          // compact CSR: filter m = taps [off[m], off[m] + count[m]) over bins start[m] ...
  short start[MAX_MELS];
  short count[MAX_MELS];
  short off[MAX_MELS];
  short total;
  double w[MAX_TOTAL_TAPS];
};

struct LogmelTables {
  double window[NFFT];
  double2 tw200[NC];     // exp(-2 pi i m / 200)
  double2 tw400[NBIN];   // exp(-2 pi i k / 400)
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 mul_neg_i(double2 a) { return make_double2(a.y, -a.x); }  // a * (-i)

__device__ __forceinline__ void dft5(double2* v) {
  const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
  const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;
  double2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
  double2 a1 = make_double2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
  double2 a2 = make_double2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
  double2 b1 = make_double2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
  double2 b2 = make_double2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
  v[0] = make_double2(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
  v[1] = make_double2(a1.x + b1.y, a1.y - b1.x);
  v[4] = make_double2(a1.x - b1.y, a1.y + b1.x);
  v[2] = make_double2(a2.x + b2.y, a2.y - b2.x);
  v[3] = make_double2(a2.x - b2.y, a2.y + b2.x);
}

__device__ __forceinline__ void dft4(double2 x0, double2 x1, double2 x2, double2 x3, double2& y0, double2& y1,
                                     double2& y2, double2& y3) {
  double2 s0 = cadd(x0, x2), s1 = csub(x0, x2), s2 = cadd(x1, x3), s3 = mul_neg_i(csub(x1, x3));
  y0 = cadd(s0, s2);
  y2 = csub(s0, s2);
  y1 = cadd(s1, s3);
  y3 = csub(s1, s3);
}

__device__ __forceinline__ void dft8(double2* v) {
  const double h = 0.70710678118654752440;
  double2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
  double2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
  b1 = make_double2(h * (b1.x + b1.y), h * (b1.y - b1.x));    // * (1 - i)/sqrt2
  b2 = mul_neg_i(b2);                                         // * (-i)
  b3 = make_double2(h * (b3.y - b3.x), h * (-b3.x - b3.y));   // * (-1 - i)/sqrt2
  dft4(a0, a1, a2, a3, v[0], v[2], v[4], v[6]);
  dft4(b0, b1, b2, b3, v[1], v[3], v[5], v[7]);
}

// One Stockham pass of radix R over a 200-point transform; Ns = product of the radices already applied.
template <int R, int Ns>
__device__ __forceinline__ void stockham_pass(const double2* __restrict__ in, double2* __restrict__ out,
                                              const double2* __restrict__ tw200, int lt) {
  constexpr int NB = NC / R;  // butterflies in this pass
  for (int j = lt; j < NB; j += TPF) {
    const int k = j % Ns;
    double2 v[R];
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] = in[j + i * NB];
    if (Ns > 1) {
      constexpr int step = NC / (Ns * R);
#pragma unroll
      for (int i = 1; i < R; ++i) v[i] = cmul(v[i], tw200[i * k * step]);
    }
    if (R == 5) dft5(v); else dft8(v);
    const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
    for (int i = 0; i < R; ++i) out[j0 + i * Ns] = v[i];
  }
}

#ifdef KW_LOGMEL_TIMING
// debug build only: per-phase cycle sums (thread 0 of every CTA); read with kw_mel_filterbank(-1, out8)
__device__ unsigned long long g_lm_dbg[8];
#define LM_T(i)                                   \
  do {                                            \
    if (threadIdx.x == 0) {                       \
      const long long now_ = clock64();           \
      dbg_acc[i] += now_ - dbg_t;                 \
      dbg_t = now_;                               \
    }                                             \
  } while (0)
#else
#define LM_T(i)
#endif

// First pass (radix 5, no twiddles) fused with the windowing: point n of the packed 200-point sequence is
// (x[2n] w[2n], x[2n+1] w[2n+1]), read straight from the staged samples instead of a separately written complex buffer.
__device__ __forceinline__ void stockham_first_pass(const float* __restrict__ x, const double* __restrict__ win,
                                                    double2* __restrict__ out, int lt) {
  constexpr int R = 5, NB = NC / R;
  for (int j = lt; j < NB; j += TPF) {
    double2 v[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int n = j + i * NB;
      const float2 s2 = *reinterpret_cast<const float2*>(x + 2 * n);
      const double2 w2 = *reinterpret_cast<const double2*>(win + 2 * n);
      v[i] = make_double2((double)s2.x * w2.x, (double)s2.y * w2.y);
    }
    dft5(v);
#pragma unroll
    for (int i = 0; i < R; ++i) out[j * R + i] = v[i];
  }
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void fill_neg_inf(float* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = -INFINITY;
}

__global__ void __launch_bounds__(NTHREADS)
logmel_stft_kernel(const float* __restrict__ audio, const long long* __restrict__ starts,
                   const int* __restrict__ lens, int n_samples, int n_frames,
                   int n_mels, const LogmelTables* __restrict__ tables, const MelBankDev* __restrict__ bank,
                   float* __restrict__ out, float* __restrict__ clip_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_window = reinterpret_cast<double*>(smem_raw);                    // 400
  double2* s_tw200 = reinterpret_cast<double2*>(s_window + NFFT);            // 200
  double2* s_tw400 = s_tw200 + NC;                                           // 201
  double2* s_fft = s_tw400 + NBIN;                                           // FPI * 2 * 200
  float* s_samp = reinterpret_cast<float*>(s_fft + FPI * 2 * NC);            // NSAMP
  float* s_tile = s_samp + NSAMP;                                            // n_mels * (FT + 1)
  __shared__ float s_wmax[NTHREADS / 32];
  // The filterbank (~400 taps, 3 KB) lives in shared memory: fetched from global / L1 inside the dependent tap loop it
  // made the mel stage 42 % of the kernel (KW_LOGMEL_TIMING build).
  __shared__ double s_bw[MAX_TOTAL_TAPS];
  __shared__ short s_bstart[MAX_MELS], s_bcount[MAX_MELS], s_boff[MAX_MELS];

  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FT;
  const int tid = threadIdx.x;
  const int len = lens ? min(lens[b], n_samples) : n_samples;
  // clip b = its own row of a [B, n_samples] matrix, or (window mode) n_samples starting at starts[b] inside one long
  // recording: samples >= len read as zero, so a 15 s window is framed in place and its pad to 30 s is synthesised here
  const float* clip = starts ? audio + starts[b] : audio + (size_t)b * n_samples;

#ifdef KW_LOGMEL_TIMING
  long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dbg_t = clock64();
#endif
  for (int i = tid; i < NFFT; i += NTHREADS) s_window[i] = tables->window[i];
  for (int i = tid; i < NC; i += NTHREADS) s_tw200[i] = tables->tw200[i];
  for (int i = tid; i < NBIN; i += NTHREADS) s_tw400[i] = tables->tw400[i];
  for (int i = tid; i < bank->total; i += NTHREADS) s_bw[i] = bank->w[i];
  for (int i = tid; i < n_mels; i += NTHREADS) {
    s_bstart[i] = bank->start[i];
    s_bcount[i] = bank->count[i];
    s_boff[i] = bank->off[i];
  }

  // stage samples [160*t0 - 200, 160*t0 - 200 + NSAMP) with reflection about 0 and n_samples-1 (no edge repeat)
  const int s_begin = HOP * t0 - NFFT / 2;
  if (s_begin >= 0 && s_begin + NSAMP <= len &&
      (reinterpret_cast<uintptr_t>(clip + s_begin) & 15) == 0) {  // interior tile inside the clip: 128-bit loads
    const float4* src = reinterpret_cast<const float4*>(clip + s_begin);  // s_begin % 4 == 0, clip 16 B aligned
    for (int i = tid; i < NSAMP / 4; i += NTHREADS) reinterpret_cast<float4*>(s_samp)[i] = __ldg(src + i);
  } else {
    for (int i = tid; i < NSAMP; i += NTHREADS) {
      int s = s_begin + i;
      if (s < 0) s = -s;
      if (s >= n_samples) s = 2 * (n_samples - 1) - s;
      s_samp[i] = (s >= 0 && s < len) ? __ldg(clip + s) : 0.0f;
    }
  }
  __syncthreads();
  LM_T(0);

  const int slot = tid / TPF, lt = tid % TPF;
  double2* bufA = s_fft + slot * 2 * NC;
  double2* bufB = bufA + NC;
  float local_max = -INFINITY;

  for (int it = 0; it < FT / FPI; ++it) {
    const int f = it * FPI + slot;  // frame within the tile
    const float* x = s_samp + HOP * f;
    stockham_first_pass(x, s_window, bufB, lt);
    __syncthreads();
    LM_T(2);
    stockham_pass<5, 5>(bufB, bufA, s_tw200, lt);
    __syncthreads();
    LM_T(3);
    stockham_pass<8, 25>(bufA, bufB, s_tw200, lt);
    __syncthreads();
    LM_T(4);
    // real-input fold: X[k] = (Z[k] + conj Z[200-k])/2 - i/2 e^{-2 pi i k/400} (Z[k] - conj Z[200-k]); power -> bufA
    double* pw = reinterpret_cast<double*>(bufA);
    for (int k = lt; k < NBIN; k += TPF) {
      double2 zk = bufB[k == NC ? 0 : k];
      double2 zc = bufB[k == 0 ? 0 : NC - k];
      zc.y = -zc.y;
      double2 e = make_double2(0.5 * (zk.x + zc.x), 0.5 * (zk.y + zc.y));
      double2 o = cmul(s_tw400[k], make_double2(0.5 * (zk.x - zc.x), 0.5 * (zk.y - zc.y)));
      double2 X = make_double2(e.x + o.y, e.y - o.x);  // e - i*o
      pw[k] = X.x * X.x + X.y * X.y;
    }
    __syncthreads();
    LM_T(5);
    for (int m = lt; m < n_mels; m += TPF) {
      const int cnt = s_bcount[m];
      const double* w = s_bw + s_boff[m];
      const double* x = pw + s_bstart[m];
      double a0 = 0.0, a1 = 0.0;  // two independent chains: the loop is latency-, not throughput-bound
      int j = 0;
      for (; j + 1 < cnt; j += 2) {
        a0 = fma(w[j], x[j], a0);
        a1 = fma(w[j + 1], x[j + 1], a1);
      }
      if (j < cnt) a0 = fma(w[j], x[j], a0);
      // lg2.approx: absolute error < 2^-22 in log2, i.e. < 2e-8 in the normalised output (bar: 1e-5)
      float lv = __log2f(fmaxf((float)(a0 + a1), 1e-10f)) * 0.30102999566398119521f;
      s_tile[m * (FT + 1) + f] = lv;
      if (t0 + f < n_frames) local_max = fmaxf(local_max, lv);
    }
    __syncthreads();
    LM_T(6);
  }

  float* dst = out + (size_t)b * n_mels * n_frames;
  for (int i = tid; i < n_mels * FT; i += NTHREADS) {
    int m = i / FT, f = i % FT;
    if (t0 + f < n_frames) dst[(size_t)m * n_frames + t0 + f] = s_tile[m * (FT + 1) + f];
  }
  local_max = warp_max(local_max);
  if ((tid & 31) == 0) s_wmax[tid >> 5] = local_max;
  __syncthreads();
  if (tid == 0) {
    float mx = s_wmax[0];
    for (int i = 1; i < NTHREADS / 32; ++i) mx = fmaxf(mx, s_wmax[i]);
    atomic_max_float(clip_max + b, mx);
  }
  LM_T(7);
#ifdef KW_LOGMEL_TIMING
  if (tid == 0)
    for (int i = 0; i < 8; ++i) atomicAdd(&g_lm_dbg[i], (unsigned long long)dbg_acc[i]);
#endif
}

#ifdef KW_LOGMEL_TIMING
void logmel_debug_dump() {
  unsigned long long h[8], z[8] = {0};
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_lm_dbg, sizeof(h));
  const char* nm[8] = {"0 tables + sample staging", "1 window -> complex", "2 radix-5 pass", "3 radix-5 pass (twiddled)",
                       "4 radix-8 pass", "5 fold + power", "6 mel + log", "7 tile store + max"};
  double tot = 0;
  for (int i = 0; i < 8; ++i) tot += (double)h[i];
  for (int i = 0; i < 8; ++i) fprintf(stderr, "  logmel phase %-28s %5.1f %%\n", nm[i], 100.0 * h[i] / tot);
  cudaMemcpyToSymbol(g_lm_dbg, z, sizeof(z));
}
#endif

// y = (max(x, clipmax - 8) + 4) / 4 in place; per_clip = n_mels * n_frames (multiple of 4)
__global__ void logmel_finalize_kernel(float* __restrict__ x, const float* __restrict__ clip_max, int per_clip4,
                                       size_t total4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < total4; i += stride) {
    const float lo = clip_max[i / per_clip4] - 8.0f;
    float4 v = reinterpret_cast<float4*>(x)[i];
    v.x = (fmaxf(v.x, lo) + 4.0f) / 4.0f;
    v.y = (fmaxf(v.y, lo) + 4.0f) / 4.0f;
    v.z = (fmaxf(v.z, lo) + 4.0f) / 4.0f;
    v.w = (fmaxf(v.w, lo) + 4.0f) / 4.0f;
    reinterpret_cast<float4*>(x)[i] = v;
  }
}

// ---- host side: filterbank + tables, built once per n_mels in float64 ------------------------------------------
static double hz_to_mel(double f) {  // Slaney scale, HF/audio_utils.py:285-296
  return f >= 1000.0 ? 15.0 + log(f / 1000.0) * (27.0 / log(6.4)) : 3.0 * f / 200.0;
}
static double mel_to_hz(double m) {  // HF/audio_utils.py:321-332
  return m >= 15.0 ? 1000.0 * exp((log(6.4) / 27.0) * (m - 15.0)) : 200.0 * m / 3.0;
}

void mel_filterbank_f64(int n_mels, std::vector<double>& fb /* [201][n_mels] */) {
  std::vector<double> ff(n_mels + 2);
  const double m0 = hz_to_mel(0.0), m1 = hz_to_mel(8000.0);
  for (int i = 0; i < n_mels + 2; ++i) {  // numpy.linspace semantics: start + i*step, last point exact
    double m = (i == n_mels + 1) ? m1 : m0 + i * ((m1 - m0) / (n_mels + 1));
    ff[i] = mel_to_hz(m);
  }
  fb.assign((size_t)NBIN * n_mels, 0.0);
  for (int k = 0; k < NBIN; ++k) {
    const double fk = (k == NBIN - 1) ? 8000.0 : k * (8000.0 / (NBIN - 1));
    for (int m = 0; m < n_mels; ++m) {
      const double down = -(ff[m] - fk) / (ff[m + 1] - ff[m]);
      const double up = (ff[m + 2] - fk) / (ff[m + 2] - ff[m + 1]);
      double v = fmax(0.0, fmin(down, up));
      fb[(size_t)k * n_mels + m] = v * (2.0 / (ff[m + 2] - ff[m]));
    }
  }
}

struct LogmelState {
  std::mutex mu;
  LogmelTables* tables = nullptr;
  MelBankDev* bank[MAX_MELS + 1] = {nullptr};
};
static LogmelState g_lm;

static int get_tables(int n_mels, const LogmelTables** tables, const MelBankDev** bank) {
  std::lock_guard<std::mutex> lock(g_lm.mu);
  if (!g_lm.tables) {
    std::vector<LogmelTables> h(1);
    const double pi = 3.14159265358979323846;
    for (int n = 0; n < NFFT; ++n) h[0].window[n] = 0.5 - 0.5 * cos(2.0 * pi * n / NFFT);
    for (int m = 0; m < NC; ++m) h[0].tw200[m] = make_double2(cos(2.0 * pi * m / NC), -sin(2.0 * pi * m / NC));
    for (int k = 0; k < NBIN; ++k) h[0].tw400[k] = make_double2(cos(2.0 * pi * k / NFFT), -sin(2.0 * pi * k / NFFT));
    KW_CUDA_OK(cudaMalloc(&g_lm.tables, sizeof(LogmelTables)));
    KW_CUDA_OK(cudaMemcpy(g_lm.tables, h.data(), sizeof(LogmelTables), cudaMemcpyHostToDevice));
  }
  if (!g_lm.bank[n_mels]) {
    std::vector<double> fb;
    mel_filterbank_f64(n_mels, fb);
    std::vector<MelBankDev> h(1);
    memset(h.data(), 0, sizeof(MelBankDev));
    int total = 0;
    for (int m = 0; m < n_mels; ++m) {
      int first = -1, last = -1;
      for (int k = 0; k < NBIN; ++k)
        if (fb[(size_t)k * n_mels + m] != 0.0) {
          if (first < 0) first = k;
          last = k;
        }
      if (first < 0) { first = 0; last = -1; }
      const int cnt = last - first + 1;
      KW_REQUIRE(total + cnt <= MAX_TOTAL_TAPS, "mel filterbank has more than %d taps", MAX_TOTAL_TAPS);
      h[0].start[m] = (short)first;
      h[0].count[m] = (short)cnt;
      h[0].off[m] = (short)total;
      for (int k = first; k <= last; ++k) h[0].w[total++] = fb[(size_t)k * n_mels + m];
    }
    h[0].total = (short)total;
    KW_CUDA_OK(cudaMalloc(&g_lm.bank[n_mels], sizeof(MelBankDev)));
    KW_CUDA_OK(cudaMemcpy(g_lm.bank[n_mels], h.data(), sizeof(MelBankDev), cudaMemcpyHostToDevice));
  }
  *tables = g_lm.tables;
  *bank = g_lm.bank[n_mels];
  return KW_OK;
}

extern std::atomic<long long> g_launches;

int logmel_launch(const float* audio, const long long* starts, const int32_t* lens, int B, int n_samples, int n_mels,
                  float* out, float* clip_max, cudaStream_t st) {
  KW_REQUIRE(B > 0 && n_mels > 0 && n_mels <= MAX_MELS, "kw_logmel: bad B=%d n_mels=%d", B, n_mels);
  KW_REQUIRE(n_samples >= NFFT, "kw_logmel: n_samples=%d must be >= 400", n_samples);
  const LogmelTables* tables;
  const MelBankDev* bank;
  int rc = get_tables(n_mels, &tables, &bank);
  if (rc) return rc;
  const int n_frames = n_samples / HOP;
  const size_t smem = sizeof(double) * NFFT + sizeof(double2) * (NC + NBIN + FPI * 2 * NC) + sizeof(float) * NSAMP +
                      sizeof(float) * n_mels * (FT + 1);
  static bool attr_set = false;
  if (!attr_set) {
    KW_CUDA_OK(cudaFuncSetAttribute(logmel_stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
    attr_set = true;
  }
  // slabs of <= 32 clips keep the un-normalised tile (<= 49 MB) resident in the 126 MB L2 for the fix-up pass
  const int SLAB = 32;
  for (int b0 = 0; b0 < B; b0 += SLAB) {
    const int nb = std::min(SLAB, B - b0);
    fill_neg_inf<<<ceil_div(nb, 128), 128, 0, st>>>(clip_max + b0, nb);
    dim3 grid(ceil_div(n_frames, FT), nb);
    logmel_stft_kernel<<<grid, NTHREADS, smem, st>>>(starts ? audio : audio + (size_t)b0 * n_samples,
                                                     starts ? starts + b0 : nullptr, lens ? lens + b0 : nullptr,
                                                     n_samples, n_frames, n_mels, tables, bank,
                                                     out + (size_t)b0 * n_mels * n_frames, clip_max + b0);
    KW_LAUNCH_OK();
    const int per_clip = n_mels * n_frames;
    KW_REQUIRE(per_clip % 4 == 0, "kw_logmel: n_mels*n_frames must be a multiple of 4");
    const size_t total4 = (size_t)nb * per_clip / 4;
    const int blocks = (int)std::min<size_t>((total4 + 255) / 256, 148 * 16);
    logmel_finalize_kernel<<<blocks, 256, 0, st>>>(out + (size_t)b0 * per_clip, clip_max + b0, per_clip / 4, total4);
    KW_LAUNCH_OK();
    g_launches += 3;
  }
  return KW_OK;
}

}  // namespace kw
