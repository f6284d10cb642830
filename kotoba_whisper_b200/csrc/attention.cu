// Attention kernels, head dim 64, fp32 arithmetic on fp32 / bf16 storage:
//   * attn_simt_kernel        encoder self-attention (and the generic plug-in seam kw_attention): flash-style, online
//                             softmax, no mask, q pre-scaled (modeling_whisper.py:310, 342-352)
//   * dec_self_attn_kernel    one decoder position: append k,v to the preallocated self-KV pool, attend over 0..pos
//                             (replaces DynamicLayer.update = torch.cat, HF/cache_utils.py:102-120)
//   * dec_cross_attn_kernel   one decoder position over the 1500 cached encoder K/V rows (modeling_whisper.py:315-336);
//                             this is the HBM-dominant kernel of a decode step
#include <atomic>

#include "common.cuh"

namespace kw {

extern std::atomic<long long> g_launches;

constexpr int HD = 64;

// accurate expf for fp32 storage (exact-mode parity), SFU ex2 for bf16 storage
template <typename T> __device__ __forceinline__ float exp_t(float x);
template <> __device__ __forceinline__ float exp_t<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ float exp_t<bf16>(float x) { return __expf(x); }

// ---------------------------------------------------------------------------------------------------------------------
// Encoder attention.  CTA = (64-query tile, head, batch), 256 threads as 16 x 16: thread (ty, tx) owns queries
// ty + 16 i and keys tx + 16 j (i, j < 4) of the 64 x 64 score tile, and output columns 4 tx .. 4 tx + 3.  The 16
// threads sharing a query row sit in one half-warp, so the row max / sum are 4 shuffles and P only needs __syncwarp.
constexpr int AT_BQ = 64, AT_BK = 64, AT_LD = HD + 4;  // +4 floats: rows stay 16 B aligned, LDS.128 conflict-free

template <typename T>
__global__ void __launch_bounds__(256)
attn_simt_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ out, int Tq,
                 int Tk, long long q_sb, long long q_st, long long kv_sb, long long kv_st, long long o_sb,
                 long long o_st, int causal) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;                      // [64][68]
  float* Ks = Qs + AT_BQ * AT_LD;      // [64][68]
  float* Vs = Ks + AT_BK * AT_LD;      // [64][68]
  float* Ps = Vs + AT_BK * AT_LD;      // [64][68]

  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
  const T* qb = q + (size_t)b * q_sb + (size_t)h * HD;
  const T* kb = k + (size_t)b * kv_sb + (size_t)h * HD;
  const T* vb = v + (size_t)b * kv_sb + (size_t)h * HD;

  // 64 rows x 16 float4 per tile -> 4 float4 per thread
  for (int i = tid; i < AT_BQ * (HD / 4); i += 256) {
    int r = i / (HD / 4), c = (i % (HD / 4)) * 4;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < Tq) val = ld4(qb + (size_t)(q0 + r) * q_st + c);
    *reinterpret_cast<float4*>(Qs + r * AT_LD + c) = val;
  }

  float m_run[4], l_run[4];
  float4 o_acc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.0f;
    o_acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  // causal (teacher-forcing decoder self-attention, create_causal_mask at modeling_whisper.py:766-772): query r sees
  // keys 0..r; key tiles entirely above the diagonal of this query tile are skipped
  const int k_end = causal ? min(Tk, q0 + AT_BQ) : Tk;
  for (int k0 = 0; k0 < k_end; k0 += AT_BK) {
    __syncthreads();  // previous tile's Ks/Vs/Ps fully consumed (also covers the Qs fill)
    for (int i = tid; i < AT_BK * (HD / 4); i += 256) {
      int r = i / (HD / 4), c = (i % (HD / 4)) * 4;
      float4 kv4 = make_float4(0.f, 0.f, 0.f, 0.f), vv4 = kv4;
      if (k0 + r < Tk) {
        kv4 = ld4(kb + (size_t)(k0 + r) * kv_st + c);
        vv4 = ld4(vb + (size_t)(k0 + r) * kv_st + c);
      }
      *reinterpret_cast<float4*>(Ks + r * AT_LD + c) = kv4;
      *reinterpret_cast<float4*>(Vs + r * AT_LD + c) = vv4;
    }
    __syncthreads();

    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.0f;
#pragma unroll 4
    for (int d = 0; d < HD; d += 4) {
      float4 qa[4], ka[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qa[i] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * i) * AT_LD + d);
#pragma unroll
      for (int j = 0; j < 4; ++j) ka[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * AT_LD + d);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = fmaf(qa[i].x, ka[j].x, s[i][j]);
          s[i][j] = fmaf(qa[i].y, ka[j].y, s[i][j]);
          s[i][j] = fmaf(qa[i].z, ka[j].z, s[i][j]);
          s[i][j] = fmaf(qa[i].w, ka[j].w, s[i][j]);
        }
    }
    // online softmax per query row
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (k0 + tx + 16 * j >= Tk || (causal && k0 + tx + 16 * j > q0 + ty + 16 * i)) s[i][j] = -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_new = fmaxf(m_run[i], mx);  // finite: tile 0 holds key 0, visible to every query
      const float corr = exp_t<T>(m_run[i] - m_new);
      float rs = 0.0f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float p = exp_t<T>(s[i][j] - m_new);
        rs += p;
        Ps[(ty + 16 * i) * AT_LD + tx + 16 * j] = p;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
      l_run[i] = l_run[i] * corr + rs;
      m_run[i] = m_new;
      o_acc[i].x *= corr; o_acc[i].y *= corr; o_acc[i].z *= corr; o_acc[i].w *= corr;
    }
    __syncwarp();  // P rows of this half-warp are written and read by the same half-warp
#pragma unroll 4
    for (int kk = 0; kk < AT_BK; kk += 4) {
      float4 pa[4], va[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pa[i] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * i) * AT_LD + kk);
#pragma unroll
      for (int j = 0; j < 4; ++j) va[j] = *reinterpret_cast<const float4*>(Vs + (kk + j) * AT_LD + 4 * tx);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o_acc[i].x = fmaf(pa[i].x, va[0].x, o_acc[i].x); o_acc[i].y = fmaf(pa[i].x, va[0].y, o_acc[i].y);
        o_acc[i].z = fmaf(pa[i].x, va[0].z, o_acc[i].z); o_acc[i].w = fmaf(pa[i].x, va[0].w, o_acc[i].w);
        o_acc[i].x = fmaf(pa[i].y, va[1].x, o_acc[i].x); o_acc[i].y = fmaf(pa[i].y, va[1].y, o_acc[i].y);
        o_acc[i].z = fmaf(pa[i].y, va[1].z, o_acc[i].z); o_acc[i].w = fmaf(pa[i].y, va[1].w, o_acc[i].w);
        o_acc[i].x = fmaf(pa[i].z, va[2].x, o_acc[i].x); o_acc[i].y = fmaf(pa[i].z, va[2].y, o_acc[i].y);
        o_acc[i].z = fmaf(pa[i].z, va[2].z, o_acc[i].z); o_acc[i].w = fmaf(pa[i].z, va[2].w, o_acc[i].w);
        o_acc[i].x = fmaf(pa[i].w, va[3].x, o_acc[i].x); o_acc[i].y = fmaf(pa[i].w, va[3].y, o_acc[i].y);
        o_acc[i].z = fmaf(pa[i].w, va[3].z, o_acc[i].z); o_acc[i].w = fmaf(pa[i].w, va[3].w, o_acc[i].w);
      }
    }
  }

  T* ob = out + (size_t)b * o_sb + (size_t)h * HD;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = q0 + ty + 16 * i;
    if (r < Tq) {
      const float inv = 1.0f / l_run[i];
      st4(ob + (size_t)r * o_st + 4 * tx,
          make_float4(o_acc[i].x * inv, o_acc[i].y * inv, o_acc[i].z * inv, o_acc[i].w * inv));
    }
  }
}

int attention_simt(const void* q, const void* k, const void* v, void* out, int B, int H, int Tq, int Tk,
                   long long q_sb, long long q_st, long long kv_sb, long long kv_st, long long o_sb, long long o_st,
                   kw_dtype t, cudaStream_t st, int causal) {
  KW_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tk > 0, "attention: empty problem");
  KW_REQUIRE(q_st % 4 == 0 && kv_st % 4 == 0 && o_st % 4 == 0 && q_sb % 4 == 0 && kv_sb % 4 == 0 && o_sb % 4 == 0,
             "attention: strides must be multiples of 4 elements");
  const size_t smem = sizeof(float) * 4 * AT_BQ * AT_LD;
  static bool attr = false;
  if (!attr) {
    KW_CUDA_OK(cudaFuncSetAttribute(attn_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KW_CUDA_OK(cudaFuncSetAttribute(attn_simt_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 grid(ceil_div(Tq, AT_BQ), H, B);
  if (t == KW_BF16)
    attn_simt_kernel<bf16><<<grid, 256, smem, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (bf16*)out, Tq, Tk,
                                                    q_sb, q_st, kv_sb, kv_st, o_sb, o_st, causal);
  else
    attn_simt_kernel<float><<<grid, 256, smem, st>>>((const float*)q, (const float*)k, (const float*)v, (float*)out,
                                                     Tq, Tk, q_sb, q_st, kv_sb, kv_st, o_sb, o_st, causal);
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

__device__ __forceinline__ void load8(const float* p, float* f) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float* f) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h2[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// 8 consecutive elements kept in their storage format until use (4 registers for bf16) so XA_UNROLL rows fit in flight
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
  uint4 r;
  __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void load_hint(const bf16* p, uint64_t pol) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  }
  __device__ __forceinline__ void unpack(float* f) const {
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h2[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void load_hint(const float* p, uint64_t pol) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p), "l"(pol));
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 4), "l"(pol));
  }
  __device__ __forceinline__ void unpack(float* f) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// Decoder self-attention, one position.  qkv f32 [B, 3d] (q pre-scaled | k | v) from the fused projection;
// pools kc/vc typed [B, H, max_t, 64]; out typed [B, d] (the next projection's operand).  CTA = (head, batch), 128 threads.
template <typename T>
__global__ void __launch_bounds__(128)
dec_self_attn_kernel(const float* __restrict__ qkv, T* __restrict__ kc, T* __restrict__ vc, T* __restrict__ out,
                     int d, int H, int max_t, int pos) {
  // 8 lanes share one cached key / value row (16-byte loads, coalesced 128 B per row), 4 rows per warp per load,
  // 2 loads in flight per lane: the whole <= 448-row history is ~14 dependent-free iterations instead of a per-thread
  // walk over 64-element rows.
  __shared__ float s_p[512];
  __shared__ float s_red[4];
  __shared__ float s_o[4][HD];
  pdl_trigger();
  pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, sub = lane >> 3, l8 = lane & 7;
  const float* row = qkv + (size_t)b * 3 * d;
  T* kp = kc + ((size_t)b * H + h) * max_t * HD;
  T* vp = vc + ((size_t)b * H + h) * max_t * HD;
  if (tid < HD) st_f(kp + (size_t)pos * HD + tid, row[d + h * HD + tid]);
  else st_f(vp + (size_t)pos * HD + (tid - HD), row[2 * d + h * HD + (tid - HD)]);
  float qf[8];
  load8(row + h * HD + l8 * 8, qf);
  __syncthreads();  // the new k/v row (written by this CTA) is visible to the whole CTA
  const int n = pos + 1;
  float lmax = -INFINITY;
  for (int j0 = 0; j0 < n; j0 += 32) {
    Raw8<T> kr[2];
    int jj[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      jj[u] = j0 + (u * 4 + warp) * 4 + sub;
      if (jj[u] < n) kr[u].load(kp + (size_t)jj[u] * HD + l8 * 8);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float acc = 0.0f;
      if (jj[u] < n) {
        float kf[8];
        kr[u].unpack(kf);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc = fmaf(qf[e], kf[e], acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (jj[u] < n) {
        if (l8 == 0) s_p[jj[u]] = acc;
        lmax = fmaxf(lmax, acc);
      }
    }
  }
  lmax = warp_max(lmax);
  if (lane == 0) s_red[warp] = lmax;
  __syncthreads();
  const float mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  __syncthreads();
  float lsum = 0.0f;
  for (int j = tid; j < n; j += 128) {
    float p = exp_t<T>(s_p[j] - mx);
    s_p[j] = p;
    lsum += p;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) s_red[warp] = lsum;
  __syncthreads();
  const float inv = 1.0f / (s_red[0] + s_red[1] + s_red[2] + s_red[3]);
  float o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.0f;
  for (int j0 = 0; j0 < n; j0 += 32) {
    Raw8<T> vr[2];
    int jj[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      jj[u] = j0 + (u * 4 + warp) * 4 + sub;
      if (jj[u] < n) vr[u].load(vp + (size_t)jj[u] * HD + l8 * 8);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (jj[u] < n) {
        const float p = s_p[jj[u]];
        float vf[8];
        vr[u].unpack(vf);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(p, vf[e], o[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    o[e] += __shfl_xor_sync(0xffffffffu, o[e], 8);
    o[e] += __shfl_xor_sync(0xffffffffu, o[e], 16);
  }
  if (sub == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) s_o[warp][l8 * 8 + e] = o[e];
  }
  __syncthreads();
  if (tid < HD) st_f(out + (size_t)b * d + h * HD + tid, (s_o[0][tid] + s_o[1][tid] + s_o[2][tid] + s_o[3][tid]) * inv);
}

int dec_self_attn(const float* qkv, void* kc, void* vc, void* out, int B, int d, int H, int max_t, int pos, kw_dtype t,
                  cudaStream_t st) {
  KW_REQUIRE(pos >= 0 && pos < max_t && max_t <= 512, "dec_self_attn: pos=%d max_t=%d", pos, max_t);
  dim3 grid(H, B);
  if (t == KW_BF16)
    KW_CUDA_OK(launch_pdl(PDL_SELF_ATTN, dec_self_attn_kernel<bf16>, grid, dim3(128), 0, st, qkv, (bf16*)kc, (bf16*)vc, (bf16*)out, d, H, max_t, pos));
  else
    KW_CUDA_OK(launch_pdl(PDL_SELF_ATTN, dec_self_attn_kernel<float>, grid, dim3(128), 0, st, qkv, (float*)kc, (float*)vc, (float*)out, d, H, max_t, pos));
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Decoder cross-attention, one position.  q f32 [B, d] (pre-scaled); xkv typed [B*S, 2d] = [K | V] rows of the one-shot
// projection (row stride 2d); out typed [B, d].  CTA = (head, batch), 128 threads.  8 lanes share one key row with one
// 16 B (bf16) / two 16 B (f32) loads each, so every K/V byte is fetched exactly once with full 32 B sectors.
constexpr int XA_THREADS = 128, XA_WARPS = XA_THREADS / 32, XA_UNROLL = 4;

template <typename T>
__global__ void __launch_bounds__(XA_THREADS, 10)
dec_cross_attn_kernel(const float* __restrict__ q, const T* __restrict__ xkv, T* __restrict__ out, int d, int S,
                      int hint, int head_rows) {
  // L2 eviction priority of the K/V stream (hint != 0): rows >= head_rows are read once per position and are far larger
  // than L2 -> evict_first, so they do not push out what is worth keeping (weights, the head rows).  Rows < head_rows:
  // hint 2 = evict_last (they stay resident from one position to the next), hint 1 / 3 = normal priority.
  const uint64_t pol_tail = l2_evict_first();
  uint64_t pol_head;
  if (hint == 2) pol_head = l2_evict_last();
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_head));
  // 128-thread CTAs: all B*H CTAs (1280 at B = 64) are resident at once (<= 16 per SM), so there is no partial second
  // wave; each warp keeps XA_UNROLL independent 16-byte loads per lane in flight.
  extern __shared__ float s_p[];  // [S]
  __shared__ float s_red[XA_WARPS];
  __shared__ float s_o[XA_WARPS][HD];
  pdl_trigger();
  pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, sub = lane >> 3, l8 = lane & 7;
  const size_t ld = 2 * (size_t)d;
  const T* kbase = xkv + (size_t)b * S * ld + (size_t)h * HD + l8 * 8;
  const T* vbase = kbase + d;
  float qf[8];
  load8(q + (size_t)b * d + h * HD + l8 * 8, qf);
  constexpr int KEYS_PER_ITER = XA_WARPS * 4 * XA_UNROLL;  // 64 keys per CTA iteration

  float lmax = -INFINITY;
  for (int j0 = 0; j0 < S; j0 += KEYS_PER_ITER) {
    Raw8<T> kr[XA_UNROLL];
    int jj[XA_UNROLL];
#pragma unroll
    for (int u = 0; u < XA_UNROLL; ++u) {
      jj[u] = j0 + (u * XA_WARPS + warp) * 4 + sub;
      if (jj[u] < S) {
        if (hint) kr[u].load_hint(kbase + (size_t)jj[u] * ld, jj[u] < head_rows ? pol_head : pol_tail);
        else kr[u].load(kbase + (size_t)jj[u] * ld);
      }
    }
#pragma unroll
    for (int u = 0; u < XA_UNROLL; ++u) {
      float acc = 0.0f;
      if (jj[u] < S) {
        float kf[8];
        kr[u].unpack(kf);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc = fmaf(qf[e], kf[e], acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (jj[u] < S) {
        if (l8 == 0) s_p[jj[u]] = acc;
        lmax = fmaxf(lmax, acc);
      }
    }
  }
  lmax = warp_max(lmax);
  if (lane == 0) s_red[warp] = lmax;
  __syncthreads();
  float mx = s_red[0];
#pragma unroll
  for (int i = 1; i < XA_WARPS; ++i) mx = fmaxf(mx, s_red[i]);
  __syncthreads();
  float lsum = 0.0f;
  for (int j = tid; j < S; j += XA_THREADS) {
    float p = exp_t<T>(s_p[j] - mx);
    s_p[j] = p;
    lsum += p;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) s_red[warp] = lsum;
  __syncthreads();
  float tot = 0.0f;
#pragma unroll
  for (int i = 0; i < XA_WARPS; ++i) tot += s_red[i];
  const float inv = 1.0f / tot;

  float o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.0f;
  for (int j0 = 0; j0 < S; j0 += KEYS_PER_ITER) {
    Raw8<T> vr[XA_UNROLL];
    int jj[XA_UNROLL];
#pragma unroll
    for (int u = 0; u < XA_UNROLL; ++u) {
      jj[u] = j0 + (u * XA_WARPS + warp) * 4 + sub;
      if (jj[u] < S) {
        if (hint) vr[u].load_hint(vbase + (size_t)jj[u] * ld, jj[u] < head_rows ? pol_head : pol_tail);
        else vr[u].load(vbase + (size_t)jj[u] * ld);
      }
    }
#pragma unroll
    for (int u = 0; u < XA_UNROLL; ++u) {
      if (jj[u] < S) {
        const float p = s_p[jj[u]];
        float vf[8];
        vr[u].unpack(vf);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(p, vf[e], o[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    o[e] += __shfl_xor_sync(0xffffffffu, o[e], 8);
    o[e] += __shfl_xor_sync(0xffffffffu, o[e], 16);
  }
  if (sub == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) s_o[warp][l8 * 8 + e] = o[e];
  }
  __syncthreads();
  if (tid < HD) {
    float acc = 0.0f;
#pragma unroll
    for (int w = 0; w < XA_WARPS; ++w) acc += s_o[w][tid];
    st_f(out + (size_t)b * d + h * HD + tid, acc * inv);
  }
}

int dec_cross_attn(const float* q, const void* xkv, void* out, int B, int d, int H, int S, kw_dtype t,
                   cudaStream_t st, int hint, int head_rows) {
  dim3 grid(H, B);
  const size_t smem = sizeof(float) * S;
  if (t == KW_BF16)
    KW_CUDA_OK(launch_pdl(PDL_CROSS_ATTN, dec_cross_attn_kernel<bf16>, grid, dim3(XA_THREADS), smem, st, q, (const bf16*)xkv, (bf16*)out, d, S, hint, head_rows));
  else
    KW_CUDA_OK(launch_pdl(PDL_CROSS_ATTN, dec_cross_attn_kernel<float>, grid, dim3(XA_THREADS), smem, st, q, (const float*)xkv, (float*)out, d, S, hint, head_rows));
  KW_LAUNCH_OK();
  ++g_launches;
  return KW_OK;
}

}  // namespace kw
