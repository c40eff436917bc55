// CUDA-core (fp32 FMA) kernels: the generic implicit-GEMM conv used by the
// low-FLOP layers (TextEncoder, flow; 2.5 % of the path's FLOPs) and as the
// validation twin of the tcgen05 conv, plus the small fused elementwise /
// normalisation / attention kernels of the TextEncoder and the flow.
//
// Reference semantics: rvc/lib/algorithm/{encoders,attentions,normalization,
// residuals,modules,commons}.py -- cited per kernel.
#include <curand_kernel.h>

#include "../../include/polgen_rvc.h"
#include "pg_common.cuh"

namespace pg {

// ---------------------------------------------------------------------------
// generic conv1d / GEMM, time-major.  Tile 64 (time) x 64 (cout), K step 16.
// ---------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void load4<__half>(const __half* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __half2 a = *reinterpret_cast<__half2*>(&t.x), b = *reinterpret_cast<__half2*>(&t.y);
  float2 fa = __half22float2(a), fb = __half22float2(b);
  v[0] = fa.x; v[1] = fa.y; v[2] = fb.x; v[3] = fb.y;
}
template <typename T> __device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store4<__half>(__half* p, const float (&v)[4]) {
  __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

constexpr int SBM = 64, SBN = 64, SBK = 16;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvArgs a) {
  __shared__ __align__(16) float As[SBK][SBM + 4];
  __shared__ __align__(16) float Bs[SBK][SBN];
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * SBM;
  const int n0 = blockIdx.y * SBN;
  const int ty = tid >> 4, tx = tid & 15;
  const TIn* x = reinterpret_cast<const TIn*>(a.x);
  const int len = a.lens ? a.lens[b] : a.L_in;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // A-load mapping: 4 consecutive channels of one row
  const int am = tid >> 2, ak = (tid & 3) * 4;
  // B-load mapping: 4 consecutive couts of one k row
  const int bk = tid >> 4, bn = (tid & 15) * 4;

  for (int tap = 0; tap < a.K; ++tap) {
    const int tin = t0 + am + tap * a.dil - a.pad;
    const bool row_ok = tin >= 0 && tin < a.L_in && (!a.in_mask || tin < len);
    const TIn* xrow = x + ((size_t)b * a.L_in + (row_ok ? tin : 0)) * a.x_ld + a.x_coff;
    const float* wtap = a.w + (size_t)tap * a.Cin * a.Cout;
    for (int c0 = 0; c0 < a.Cin; c0 += SBK) {
      float av[4] = {0.f, 0.f, 0.f, 0.f};
      if (row_ok && c0 + ak < a.Cin) {   // Cin is a multiple of 4 (checked by the launcher)
        load4<TIn>(xrow + c0 + ak, av);
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = av[i] > 0.f ? av[i] : av[i] * a.in_slope;
      }
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + bk < a.Cin && n0 + bn < a.Cout)
        bv = *reinterpret_cast<const float4*>(wtap + (size_t)(c0 + bk) * a.Cout + n0 + bn);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) As[ak + i][am] = av[i];
      *reinterpret_cast<float4*>(&Bs[bk][bn]) = bv;
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < SBK; ++kk) {
        float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
        const float br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
    }
  }

  // epilogue
  const int n = n0 + tx * 4;
  if (n >= a.Cout) return;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) bias[j] = a.bias[n + j];
  }
  if (a.bbias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) bias[j] += a.bbias[(size_t)b * a.bbias_ld + n + j];
  }
  TOut* y = reinterpret_cast<TOut*>(a.y);
  const TOut* res = reinterpret_cast<const TOut*>(a.res);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty * 4 + i;
    if (t >= a.L_out) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
    const size_t row = (size_t)b * a.L_out + t;
    if (res) {
      float r[4];
      load4<TOut>(res + row * a.res_ld + a.res_coff + n, r);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += a.res_scale * r[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] *= a.out_scale;
    TOut* yp = y + row * a.y_ld + a.y_coff + n;
    if (a.accumulate) {
      float o[4];
      load4<TOut>(yp, o);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += o[j];
    }
    if (a.act == ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * a.out_slope;
    } else if (a.act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (a.act == ACT_TANH) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = tanhf(v[j]);
    }
    if (a.out_mask && t >= len) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = 0.f;
    }
    store4<TOut>(yp, v);
  }
}

cudaError_t launch_conv_simt(const ConvArgs& a, DType in_dt, DType out_dt, cudaStream_t s) {
  if (a.Cin % 4 || a.Cout % 4 || a.x_coff % 4 || a.x_ld % 4 || a.y_ld % 4 || a.y_coff % 4 ||
      (a.res && (a.res_ld % 4 || a.res_coff % 4)))
    return cudaErrorInvalidValue;
  dim3 grid((a.L_out + SBM - 1) / SBM, (a.Cout + SBN - 1) / SBN, a.B);
  if (in_dt == DT_F32 && out_dt == DT_F32)
    conv_simt_kernel<float, float><<<grid, 256, 0, s>>>(a);
  else if (in_dt == DT_F32 && out_dt == DT_F16)
    conv_simt_kernel<float, __half><<<grid, 256, 0, s>>>(a);
  else if (in_dt == DT_F16 && out_dt == DT_F16)
    conv_simt_kernel<__half, __half><<<grid, 256, 0, s>>>(a);
  else
    conv_simt_kernel<__half, float><<<grid, 256, 0, s>>>(a);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
__global__ void set_call_kernel(CallMeta* meta, uint64_t* seeds, int B, int t_true, int ragged, uint64_t seed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    meta->t_true = t_true;
    meta->ragged = ragged;
  }
  if (i < B) seeds[i] = row_seed(seed, i);
}

cudaError_t launch_set_call(CallMeta* meta, uint64_t* seeds, int B, int t_true, int ragged, uint64_t seed,
                            cudaStream_t s) {
  set_call_kernel<<<(B + 255) / 256, 256, 0, s>>>(meta, seeds, B, t_true, ragged, seed);
  return cudaGetLastError();
}

__global__ void set_rows_kernel(int64_t* lengths, int64_t* sid, const RowScalars rs, int n) {
  const int i = threadIdx.x;
  if (i < n) {
    lengths[i] = rs.len[i];
    sid[i] = rs.sid[i];
  }
}

cudaError_t launch_set_rows(int64_t* lengths, int64_t* sid, const RowScalars& rs, int n, cudaStream_t s) {
  set_rows_kernel<<<1, RowScalars::N, 0, s>>>(lengths, sid, rs, n);
  return cudaGetLastError();
}

__global__ void prepare_ints_kernel(const int64_t* lengths, const int64_t* pitch, const int64_t* sid,
                                    const CallMeta* meta, int* lens32, int* tlen32, int* pitch32, int* sid32,
                                    int B, int T, int n_spk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int t_true = meta ? min(meta->t_true, T) : T;
  const int ragged = meta ? meta->ragged : 0;
  if (i < B) {
    int len = t_true;
    if (lengths) {
      long long l = lengths[i];
      len = l < 0 ? 0 : (l > t_true ? t_true : (int)l);
      lens32[i] = len;
    }
    if (tlen32) tlen32[i] = ragged ? len : t_true;
    if (sid) {
      long long v = sid[i];
      sid32[i] = v < 0 ? 0 : (v >= n_spk ? n_spk - 1 : (int)v);
    }
  }
  if (pitch && i < B * T) {
    long long p = pitch[i];
    pitch32[i] = p < 0 ? 0 : (p > 255 ? 255 : (int)p);
  }
}

cudaError_t launch_prepare_ints(const int64_t* lengths, const int64_t* pitch, const int64_t* sid,
                                const CallMeta* meta, int* lens32, int* tlen32, int* pitch32, int* sid32,
                                int B, int T, int n_spk, cudaStream_t s) {
  const int n = B * T > B ? B * T : B;
  prepare_ints_kernel<<<(n + 255) / 256, 256, 0, s>>>(lengths, pitch, sid, meta, lens32, tlen32, pitch32,
                                                      sid32, B, T, n_spk);
  return cudaGetLastError();
}

// emb_g lookup (synthesizers.py:172) + 1x1 conv on a length-1 sequence
// (nsf.py:125-126 `cond`, modules.py:63-64 `cond_layer`): one warp per output.
__global__ void cond_gemv_kernel(const float* __restrict__ emb, const int* __restrict__ sid,
                                 const float* __restrict__ w, const float* __restrict__ bias,
                                 float* __restrict__ y, int B, int Kdim, int N) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * N) return;
  const int b = warp / N, n = warp % N;
  const float* e = emb + (size_t)sid[b] * Kdim;
  const float* wr = w + (size_t)n * Kdim;
  float s = 0.f;
  for (int k = lane; k < Kdim; k += 32) s = fmaf(wr[k], e[k], s);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[(size_t)b * N + n] = s + (bias ? bias[n] : 0.f);
}

cudaError_t launch_cond_gemv(const float* emb, const int* sid, const float* w, const float* bias,
                             float* y, int B, int Kdim, int N, cudaStream_t s) {
  const int warps = B * N;
  cond_gemv_kernel<<<(warps * 32 + 255) / 256, 256, 0, s>>>(emb, sid, w, bias, y, B, Kdim, N);
  return cudaGetLastError();
}

// encoders.py:114-122: (emb_phone(phone) + emb_pitch(pitch)) * sqrt(H) -> LeakyReLU(0.1) -> * mask
__global__ void embed_finish_kernel(float* __restrict__ x, const float* __restrict__ emb_pitch,
                                    const int* __restrict__ pitch, const int* __restrict__ lens,
                                    int B, int T, int H, float scale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * T * H;
  if (i >= total) return;
  const int c = (int)(i % H);
  const size_t row = i / H;
  const int t = (int)(row % T), b = (int)(row / T);
  float v = x[i];
  if (pitch) v += emb_pitch[(size_t)pitch[row] * H + c];
  v *= scale;
  v = v > 0.f ? v : 0.1f * v;
  x[i] = t < lens[b] ? v : 0.f;
}

cudaError_t launch_embed_finish(float* x, const float* emb_pitch, const int* pitch, const int* lens,
                                int B, int T, int H, float scale, cudaStream_t s) {
  const size_t total = (size_t)B * T * H;
  embed_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, emb_pitch, pitch, lens, B, T,
                                                                      H, scale);
  return cudaGetLastError();
}

// normalization.py:13-16 applied to (x + y) (encoders.py:66,70): one warp per row.
template <int PER_LANE>
__global__ void add_layernorm_kernel(float* __restrict__ x, const float* __restrict__ y,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     int rows, int H) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[PER_LANE];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < H ? x[(size_t)row * H + c] + y[(size_t)row * H + c] : 0.f;
    s += v[i];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s / H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const int c = lane + 32 * i;
    const float d = c < H ? v[i] - mu : 0.f;
    q += d * d;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / H + 1e-5f);
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const int c = lane + 32 * i;
    if (c < H) x[(size_t)row * H + c] = (v[i] - mu) * rstd * gamma[c] + beta[c];
  }
}

cudaError_t launch_add_layernorm(float* x, const float* y, const float* gamma, const float* beta,
                                 int rows, int H, cudaStream_t s) {
  if (H > 256) return cudaErrorInvalidValue;
  add_layernorm_kernel<8><<<(rows * 32 + 255) / 256, 256, 0, s>>>(x, y, gamma, beta, rows, H);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// windowed relative-position attention (attentions.py:63-113, helpers :115-158)
//   scores[i,j] = q_i.k_j/sqrt(d) + [|j-i|<=w] q_i.Ek[j-i+w]/sqrt(d)
//   masked_fill(mask_i*mask_j==0, -1e4); softmax_j
//   out_i = sum_j p_ij v_j + sum_{|j-i|<=w} p_ij Ev[j-i+w]
// flash-style: 64 queries x 64 keys per step, online softmax, fp32 throughout.
// ---------------------------------------------------------------------------
constexpr int ABQ = 64, ABK = 64, AMAXR = 32;

template <int D>
__global__ void __launch_bounds__(256) rel_attention_kernel(
    const float* __restrict__ qkv, const float* __restrict__ rel_k, const float* __restrict__ rel_v,
    const int* __restrict__ lens, float* __restrict__ out, int T, int H, int window) {
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                       // [D][ABQ+4]
  float* Ks = Qs + D * (ABQ + 4);         // [D][ABK+4]
  float* Vs = Ks + D * (ABK + 4);         // [ABK][D]
  float* Ps = Vs + ABK * D;               // [ABQ][ABK+1]
  float* RL = Ps + ABQ * (ABK + 1);       // [ABQ][AMAXR]  rel-key logits
  float* PR = RL + ABQ * AMAXR;           // [ABQ][AMAXR]  band probabilities

  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ABQ;
  const int len = lens[b];
  const int R = 2 * window + 1;
  const int ld = 3 * H;
  const float* base = qkv + (size_t)b * T * ld;
  const float qscale = rsqrtf((float)D);
  constexpr int DPT = D / 16;             // output channels per thread

  for (int idx = tid; idx < ABQ * D; idx += 256) {
    const int i = idx / D, c = idx % D;
    const int t = q0 + i;
    Qs[c * (ABQ + 4) + i] = t < T ? base[(size_t)t * ld + h * D + c] * qscale : 0.f;
  }
  for (int idx = tid; idx < ABQ * AMAXR; idx += 256) PR[idx] = 0.f;
  __syncthreads();
  for (int idx = tid; idx < ABQ * R; idx += 256) {
    const int i = idx / R, r = idx % R;
    float s = 0.f;
    for (int c = 0; c < D; ++c) s = fmaf(Qs[c * (ABQ + 4) + i], rel_k[r * D + c], s);
    RL[i * AMAXR + r] = s;
  }

  float m_run[4], l_run[4], o[4][DPT];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    m_run[a] = -INFINITY;
    l_run[a] = 0.f;
#pragma unroll
    for (int c = 0; c < DPT; ++c) o[a][c] = 0.f;
  }

  for (int k0 = 0; k0 < T; k0 += ABK) {
    __syncthreads();   // previous tile fully consumed (also orders RL/PR init)
    for (int idx = tid; idx < ABK * D; idx += 256) {
      const int j = idx / D, c = idx % D;
      const int t = k0 + j;
      float kv = 0.f, vv = 0.f;
      if (t < T) {
        kv = base[(size_t)t * ld + H + h * D + c];
        vv = base[(size_t)t * ld + 2 * H + h * D + c];
      }
      Ks[c * (ABK + 4) + j] = kv;
      Vs[j * D + c] = vv;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) s[a][bb] = 0.f;
#pragma unroll 8
    for (int c = 0; c < D; ++c) {
      const float4 q4 = *reinterpret_cast<const float4*>(&Qs[c * (ABQ + 4) + ty * 4]);
      const float4 k4 = *reinterpret_cast<const float4*>(&Ks[c * (ABK + 4) + tx * 4]);
      const float qa[4] = {q4.x, q4.y, q4.z, q4.w};
      const float ka[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) s[a][bb] = fmaf(qa[a], ka[bb], s[a][bb]);
    }
    float alpha[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = q0 + ty * 4 + a;
      float mx = -INFINITY;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int j = k0 + tx * 4 + bb;
        const int rel = j - i + window;
        float v = s[a][bb];
        if (rel >= 0 && rel < R) v += RL[(ty * 4 + a) * AMAXR + rel];
        if (!(i < len && j < len)) v = -1e4f;
        if (j >= T) v = -INFINITY;
        s[a][bb] = v;
        mx = fmaxf(mx, v);
      }
#pragma unroll
      for (int off = 8; off; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[a], mx);
      alpha[a] = __expf(m_run[a] - m_new);
      float rs = 0.f;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const float p = __expf(s[a][bb] - m_new);
        Ps[(ty * 4 + a) * (ABK + 1) + tx * 4 + bb] = p;
        rs += p;
      }
#pragma unroll
      for (int off = 8; off; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l_run[a] = l_run[a] * alpha[a] + rs;
      m_run[a] = m_new;
    }
    __syncthreads();
    // band probabilities for the relative-value term (attentions.py:106-111)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int il = ty * 4 + a;
      for (int r = tx; r < R; r += 16) {
        float pr = PR[il * AMAXR + r] * alpha[a];
        const int j = q0 + il + r - window;
        if (j >= k0 && j < k0 + ABK && j < T) pr += Ps[il * (ABK + 1) + (j - k0)];
        PR[il * AMAXR + r] = pr;
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < DPT; ++c) o[a][c] *= alpha[a];
#pragma unroll 4
    for (int j = 0; j < ABK; ++j) {
      float vv[DPT];
#pragma unroll
      for (int c = 0; c < DPT; ++c) vv[c] = Vs[j * D + tx * DPT + c];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float p = Ps[(ty * 4 + a) * (ABK + 1) + j];
#pragma unroll
        for (int c = 0; c < DPT; ++c) o[a][c] = fmaf(p, vv[c], o[a][c]);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int il = ty * 4 + a, t = q0 + il;
    if (t >= T) continue;
    const float inv = 1.f / l_run[a];
#pragma unroll
    for (int c = 0; c < DPT; ++c) {
      float v = o[a][c];
      for (int r = 0; r < R; ++r) v = fmaf(PR[il * AMAXR + r], rel_v[r * D + tx * DPT + c], v);
      out[((size_t)b * T + t) * H + h * D + tx * DPT + c] = v * inv;
    }
  }
}

cudaError_t launch_rel_attention(const float* qkv, const float* rel_k, const float* rel_v,
                                 const int* lens, float* out, int B, int T, int H, int n_heads,
                                 int window, cudaStream_t s) {
  const int D = H / n_heads;
  if (D != 96 || 2 * window + 1 > AMAXR) return cudaErrorInvalidValue;
  constexpr int DD = 96;
  const size_t smem = sizeof(float) * (DD * (ABQ + 4) + DD * (ABK + 4) + ABK * DD + ABQ * (ABK + 1) +
                                       2 * ABQ * AMAXR);
  static DeviceOnce once;
  if (cudaError_t e = ensure_dyn_smem(rel_attention_kernel<DD>, once, (int)smem)) return e;
  dim3 grid((T + ABQ - 1) / ABQ, n_heads, B);
  rel_attention_kernel<DD><<<grid, 256, smem, s>>>(qkv, rel_k, rel_v, lens, out, T, H, window);
  return cudaGetLastError();
}

// synthesizers.py:174: z_p = (m_p + exp(logs_p) * randn * 0.66666) * x_mask
__global__ void reparam_kernel(const float* __restrict__ stats, const float* __restrict__ eps, int eps_T,
                               uint64_t seed, const uint64_t* __restrict__ seeds_dev,
                               const int* __restrict__ lens, float* __restrict__ m_p,
                               float* __restrict__ logs_p, float* __restrict__ z_p,
                               float* __restrict__ z, int B, int T, int C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * T * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  const size_t row = i / C;
  const int t = (int)(row % T), b = (int)(row / T);
  const float m = stats[row * 2 * C + c], lg = stats[row * 2 * C + C + c];
  float v = 0.f;
  if (t < lens[b]) {
    float e;
    if (eps) {
      e = t < eps_T ? eps[((size_t)b * eps_T + t) * C + c] : 0.f;
    } else {
      curandStatePhilox4_32_10_t st;
      curand_init(seeds_dev ? seeds_dev[b] : row_seed(seed, b), (unsigned long long)t * C + c, 0, &st);
      e = curand_normal(&st);
    }
    v = m + __expf(lg) * e * 0.66666f;
  }
  m_p[i] = m;
  logs_p[i] = lg;
  z_p[i] = v;
  z[i] = v;
}

cudaError_t launch_reparam(const float* stats, const float* eps, int eps_T, uint64_t seed,
                           const uint64_t* seeds_dev, const int* lens, float* m_p, float* logs_p, float* z_p,
                           float* z, int B, int T, int C, cudaStream_t s) {
  const size_t total = (size_t)B * T * C;
  reparam_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(stats, eps, eps_T, seed, seeds_dev, lens, m_p,
                                                                 logs_p, z_p, z, B, T, C);
  return cudaGetLastError();
}

// commons.py:79-86 fused_add_tanh_sigmoid_multiply (the +g term is already in `a`)
__global__ void gate_kernel(const float* __restrict__ a, float* __restrict__ acts, int64_t rows, int H) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H) return;
  const int64_t row = i / H;
  const int c = (int)(i % H);
  const float t = a[row * 2 * H + c], g = a[row * 2 * H + H + c];
  acts[i] = tanhf(t) * (1.f / (1.f + __expf(-g)));
}

cudaError_t launch_gate(const float* a, float* acts, int64_t rows, int H, cudaStream_t s) {
  const int64_t total = rows * H;
  gate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a, acts, rows, H);
  return cudaGetLastError();
}

__global__ void cast_f16_to_f32_kernel(const __half* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __half2float(x[i]);
}
cudaError_t launch_cast_f16_to_f32(const __half* x, float* y, int64_t n, cudaStream_t s) {
  cast_f16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, y, n);
  return cudaGetLastError();
}

}  // namespace pg
