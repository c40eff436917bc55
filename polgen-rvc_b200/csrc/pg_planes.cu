// Bandwidth kernels of the decoder over the channel-plane layout f16/f32 [B][C/8][L][8]
// (see pg_conv_planes.cu): layout conversion at the decoder's input and at the parity taps,
// the NSF source injection (reference rvc/lib/algorithm/nsf.py:93-101, :131) and
// leaky_relu(0.01) -> conv_post -> tanh (nsf.py:142-143).
// Thread mapping everywhere: consecutive threads = consecutive time rows of one plane, so every
// plane access is a coalesced 16-byte (f16) / 32-byte (f32) per-lane vector.
#include "pg_common.cuh"

namespace pg {

namespace {

__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 q;
  __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  return q;
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
  unpack8(*reinterpret_cast<const uint4*>(p), v);
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__half* p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) = pack8(v);
}

// (b, plane, t) of a flat index with t fastest
struct Idx { int b, pl, t; };
__device__ __forceinline__ Idx split(size_t i, int L, int CP) {
  Idx r;
  r.t = (int)(i % L);
  const size_t q = i / L;
  r.pl = (int)(q % CP);
  r.b = (int)(q / CP);
  return r;
}

template <typename TIn>
__global__ void nlc_to_planes_kernel(const TIn* __restrict__ x, int x_ld, int x_coff, __half* __restrict__ y,
                                     int B, int L, int C, const int* __restrict__ lens, float slope) {
  const int CP = C / 8;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * CP * L) return;
  const Idx id = split(i, L, CP);
  float v[8];
  if (lens && id.t >= lens[id.b]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = 0.f;
  } else {
    load8(x + ((size_t)id.b * L + id.t) * x_ld + x_coff + id.pl * 8, v);
    if (slope != 1.f) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = v[k] > 0.f ? v[k] : v[k] * slope;
    }
  }
  store8(y + i * 8, v);
}

template <typename TIn, typename TOut>
__global__ void planes_to_nlc_kernel(const TIn* __restrict__ x, TOut* __restrict__ y, int B, int L, int C,
                                     float inv_slope) {
  const int CP = C / 8;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * CP * L) return;
  const Idx id = split(i, L, CP);
  float v[8];
  load8(x + i * 8, v);
  if (inv_slope != 1.f) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = v[k] < 0.f ? v[k] * inv_slope : v[k];
  }
  store8(y + ((size_t)id.b * L + id.t) * C + id.pl * 8, v);
}

// nsf.py:131  x = x + noise_convs[i](har_source), then the L-form copy the next conv reads
template <typename T>
__global__ void __launch_bounds__(256) noise_inject_planes_kernel(
    T* __restrict__ x, __half* __restrict__ a16, __half* __restrict__ lo16, const float* __restrict__ src,
    const float* __restrict__ wn, const float* __restrict__ bn, int B, int L, int C, int Lsrc, int k, int stride,
    int pad, float slope) {
  const int CP = C / 8;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * CP * L) return;
  const Idx id = split(i, L, CP);
  const int c0 = id.pl * 8;
  float acc[8];
  load8(bn + c0, acc);
  const float* sp = src + (size_t)id.b * Lsrc;
  const int s0 = id.t * stride - pad;
  for (int j = 0; j < k; ++j) {
    const int n = s0 + j;
    if (n < 0 || n >= Lsrc) continue;
    const float sv = sp[n];
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wn + (size_t)j * C + c0));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(wn + (size_t)j * C + c0 + 4));
    acc[0] = fmaf(w0.x, sv, acc[0]); acc[1] = fmaf(w0.y, sv, acc[1]);
    acc[2] = fmaf(w0.z, sv, acc[2]); acc[3] = fmaf(w0.w, sv, acc[3]);
    acc[4] = fmaf(w1.x, sv, acc[4]); acc[5] = fmaf(w1.y, sv, acc[5]);
    acc[6] = fmaf(w1.z, sv, acc[6]); acc[7] = fmaf(w1.w, sv, acc[7]);
  }
  float v[8];
  load8(x + i * 8, v);
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] += acc[q];
  if (sizeof(T) == 4 && !lo16) store8(x + i * 8, v);   // fp32 residual stream keeps the raw value
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] = v[q] > 0.f ? v[q] : v[q] * slope;
  const uint4 hi = pack8(v);
  *reinterpret_cast<uint4*>(a16 + i * 8) = hi;
  if (lo16) {                                           // hi/lo stream: the f16 rounding remainder
    float hv[8];
    unpack8(hi, hv);
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] -= hv[q];
    store8(lo16 + i * 8, v);
  }
}


// Tiled form for the strided stages (k = 2*stride taps, f16 stream): one block = 32 time rows x all C
// channels of one batch row.  The source window, the [k][C] weights and the x tile are staged in shared
// memory with coalesced loads; a lane owns one 8-channel chunk (weights: two conflict-free 16 B reads per
// tap, source sample: a broadcast) and CP/8 rows, so every weight read feeds 8*CP/8 FMAs.
constexpr int NRB = 32;   // rows per sub-tile
constexpr int NSUB = 8;   // sub-tiles per block: the [k][C] weights are staged once per 256 rows (r02: 207 -> us per launch
                          // was dominated by re-staging them for every 32 rows)
template <int CP>
__global__ void __launch_bounds__(256) noise_inject_tiled_kernel(
    __half* __restrict__ x, const float* __restrict__ src, const float* __restrict__ wn,
    const float* __restrict__ bn, int L, int Lsrc, int k, int stride, int pad, float slope) {
  constexpr int C = CP * 8;
  constexpr int NR = CP / 8;              // rows per thread
  constexpr int XP = NRB * 16 + 16;       // x-tile plane pitch in bytes (bank-staggered)
  extern __shared__ __align__(16) uint8_t nsm[];
  float* ws = reinterpret_cast<float*>(nsm);                    // [k][C]
  float* ss = ws + (size_t)k * C;                               // [(NRB-1)*stride + k]
  uint8_t* xs = reinterpret_cast<uint8_t*>(ss + (((NRB - 1) * stride + k + 3) & ~3));   // [CP][NRB] x 16 B
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  for (int i = tid; i < k * C; i += 256) ws[i] = wn[i];
  const float* sp = src + (size_t)b * Lsrc;
  for (int sub = 0; sub < NSUB; ++sub) {
  const int t0 = (blockIdx.x * NSUB + sub) * NRB;
  if (t0 >= L) break;
  __syncthreads();          // the previous sub-tile's stores have read xs / ss
  const int win = (NRB - 1) * stride + k, s_first = t0 * stride - pad;
  for (int i = tid; i < win; i += 256) {
    const int n = s_first + i;
    ss[i] = (n >= 0 && n < Lsrc) ? sp[n] : 0.f;
  }
  __half* xb = x + (size_t)b * CP * L * 8;
  for (int i = tid; i < CP * NRB; i += 256) {
    const int pl = i / NRB, r = i - pl * NRB;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (t0 + r < L) v = *reinterpret_cast<const uint4*>(xb + ((size_t)pl * L + t0 + r) * 8);
    *reinterpret_cast<uint4*>(xs + pl * XP + r * 16) = v;
  }
  __syncthreads();
  const int chunk = lane % CP, rg = lane / CP;
  const int r0 = warp * (NRB / 8) + rg * NR;
  float acc[NR][8];
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) load8(bn + chunk * 8, acc[rr]);
  for (int j = 0; j < k; ++j) {
    float w[8];
    load8(ws + (size_t)j * C + chunk * 8, w);
#pragma unroll
    for (int rr = 0; rr < NR; ++rr) {
      const float sv = ss[(r0 + rr) * stride + j];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[rr][q] = fmaf(w[q], sv, acc[rr][q]);
    }
  }
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) {
    uint8_t* cell = xs + chunk * XP + (r0 + rr) * 16;
    float v[8];
    unpack8(*reinterpret_cast<const uint4*>(cell), v);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float a = v[q] + acc[rr][q];
      v[q] = a > 0.f ? a : a * slope;
    }
    *reinterpret_cast<uint4*>(cell) = pack8(v);
  }
  __syncthreads();
  for (int i = tid; i < CP * NRB; i += 256) {
    const int pl = i / NRB, r = i - pl * NRB;
    if (t0 + r < L)
      *reinterpret_cast<uint4*>(xb + ((size_t)pl * L + t0 + r) * 8) = *reinterpret_cast<const uint4*>(xs + pl * XP + r * 16);
  }
  }   // sub-tiles
}

template <int CP>
cudaError_t launch_noise_tiled(__half* x, const float* src, const float* wn, const float* bn, int B, int L,
                               int Lsrc, int k, int stride, int pad, float slope, cudaStream_t s) {
  const size_t smem = sizeof(float) * ((size_t)k * CP * 8 + (((NRB - 1) * stride + k + 3) & ~3)) +
                      (size_t)CP * (NRB * 16 + 16);
  static DeviceOnce once;
  if (smem > 48 * 1024)
    if (cudaError_t e = ensure_dyn_smem(noise_inject_tiled_kernel<CP>, once, (int)smem)) return e;
  dim3 grid((L + NRB * NSUB - 1) / (NRB * NSUB), B);
  noise_inject_tiled_kernel<CP><<<grid, 256, smem, s>>>(x, src, wn, bn, L, Lsrc, k, stride, pad, slope);
  return cudaGetLastError();
}

// nsf.py:142-143  wave = tanh(conv_post(leaky_relu(x)))  (Cout = 1, no bias).  One block = 256 output samples.
// Row-wise form: wave[t] = sum_j d_j[t + j - K/2] with d_j[n] = sum_c w[j][c] * lrelu(x[n][c]), so a thread reads ITS
// row once from global memory (plane rows of consecutive lanes are contiguous: coalesced, no staging), applies the
// leaky-ReLU once per element and keeps K running dot products; the K partials per row go through shared memory and
// a second step adds the K shifted ones.  (The round-2 kernel staged the raw window in shared memory and re-applied
// the activation for every tap: 3x the instructions and bank-conflicted 32 B row reads -- 232 us for 393 MB.)
constexpr int CPB = 256;
template <typename T, int K>
__global__ void __launch_bounds__(CPB) conv_post_planes_kernel(const T* __restrict__ x,
                                                               const float* __restrict__ w /*[K][C]*/,
                                                               float* __restrict__ wave, int L, int C,
                                                               float in_slope, const int* __restrict__ tlen,
                                                               int len_mul) {
  extern __shared__ __align__(16) float cps[];
  const int CP = C / 8, ROWS = CPB + K - 1;
  float* ws = cps;                 // [K][C]
  float* ds = cps + K * C;         // [K][ROWS]
  const int b = blockIdx.y, t0 = blockIdx.x * CPB;
  const int t_hi = tlen ? min(L, tlen[b] * len_mul) : L;   // hard end of the row
  for (int i = threadIdx.x; i < K * C; i += CPB) ws[i] = w[i];
  __syncthreads();
  const T* xb = x + (size_t)b * CP * (size_t)L * 8;
  for (int rr = threadIdx.x; rr < ROWS; rr += CPB) {
    const int n = t0 + rr - K / 2;
    float acc[K];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = 0.f;
    if (n >= 0 && n < t_hi) {
#pragma unroll 4
      for (int pl = 0; pl < CP; ++pl) {
        float v[8];
        load8(xb + ((size_t)pl * L + n) * 8, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], v[q] * in_slope);   // slope <= 1
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const float4 w0 = *reinterpret_cast<const float4*>(ws + j * C + pl * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(ws + j * C + pl * 8 + 4);
          float a = acc[j];
          a = fmaf(w0.x, v[0], a); a = fmaf(w0.y, v[1], a); a = fmaf(w0.z, v[2], a); a = fmaf(w0.w, v[3], a);
          a = fmaf(w1.x, v[4], a); a = fmaf(w1.y, v[5], a); a = fmaf(w1.z, v[6], a); a = fmaf(w1.w, v[7], a);
          acc[j] = a;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < K; ++j) ds[j * ROWS + rr] = acc[j];
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= L) return;
  float out = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) out += ds[j * ROWS + threadIdx.x + j];
  wave[(size_t)b * L + t] = t < t_hi ? tanhf(out) : 0.f;
}

// Second half of the fused post step: the last ResBlock pair left, per row, the K_POST partial dot products of each
// 16-channel half (part[h][j][b][t], see PairConvArgs::post_part); conv_post's output is the sum of the shifted
// partials -- wave[t] = tanh(sum_h sum_j part[h][j][t + j - K/2]) -- with rows outside [0, hard end) contributing zero.
__global__ void __launch_bounds__(256) conv_post_sum_kernel(const float* __restrict__ part, float* __restrict__ wave,
                                                            int B, int L, int n_half, const int* __restrict__ tlen,
                                                            int len_mul) {
  const int b = blockIdx.y, t = blockIdx.x * 256 + threadIdx.x;
  if (t >= L) return;
  const int t_hi = tlen ? min(L, tlen[b] * len_mul) : L;
  float acc = 0.f;
  if (t < t_hi) {
    for (int h = 0; h < n_half; ++h) {
#pragma unroll
      for (int j = 0; j < K_POST; ++j) {
        const int n = t + j - K_POST / 2;
        if (n >= 0 && n < t_hi) acc += part[((size_t)(h * K_POST + j) * B + b) * L + n];
      }
    }
  }
  wave[(size_t)b * L + t] = t < t_hi ? tanhf(acc) : 0.f;
}

unsigned blocks_for(size_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

cudaError_t launch_nlc_to_planes(const void* x, DType dt, int x_ld, int x_coff, __half* y, int B, int L,
                                 int C, const int* lens, float slope, cudaStream_t s) {
  if (C % 8 || x_ld % 8 || x_coff % 8) return cudaErrorInvalidValue;
  const size_t n = (size_t)B * (C / 8) * L;
  if (dt == DT_F32)
    nlc_to_planes_kernel<float><<<blocks_for(n), 256, 0, s>>>(reinterpret_cast<const float*>(x), x_ld, x_coff,
                                                             y, B, L, C, lens, slope);
  else
    nlc_to_planes_kernel<__half><<<blocks_for(n), 256, 0, s>>>(reinterpret_cast<const __half*>(x), x_ld,
                                                              x_coff, y, B, L, C, lens, slope);
  return cudaGetLastError();
}

cudaError_t launch_planes_to_nlc(const void* x, DType dt, float* y, int B, int L, int C, float inv_slope,
                                 cudaStream_t s) {
  if (C % 8) return cudaErrorInvalidValue;
  const size_t n = (size_t)B * (C / 8) * L;
  if (dt == DT_F32)
    planes_to_nlc_kernel<float, float><<<blocks_for(n), 256, 0, s>>>(reinterpret_cast<const float*>(x), y, B, L,
                                                                    C, inv_slope);
  else
    planes_to_nlc_kernel<__half, float><<<blocks_for(n), 256, 0, s>>>(reinterpret_cast<const __half*>(x), y, B,
                                                                     L, C, inv_slope);
  return cudaGetLastError();
}

cudaError_t launch_planes_to_nlc_f16(const __half* x, __half* y, int B, int L, int C, cudaStream_t s) {
  if (C % 8) return cudaErrorInvalidValue;
  const size_t n = (size_t)B * (C / 8) * L;
  planes_to_nlc_kernel<__half, __half><<<blocks_for(n), 256, 0, s>>>(x, y, B, L, C, 1.f);
  return cudaGetLastError();
}

cudaError_t launch_noise_inject_planes(void* x, DType dt, __half* a16, __half* lo16, const float* src,
                                       const float* wn, const float* bn, int B, int L, int C, int Lsrc, int k,
                                       int stride, int pad, float slope, cudaStream_t s) {
  if (C % 8) return cudaErrorInvalidValue;
  static const int tiled_min_k = [] { const char* e = getenv("PG_NOISE_TILED_MINK"); return e ? atoi(e) : 8; }();
  if (dt == DT_F16 && a16 == x && k >= tiled_min_k) {   // long-tap stages of the f16 stream: shared-memory tiled kernel
    // (measured: 268 -> 90 us at k = 80; at k = 4 the plain kernel is already HBM-bound and faster)
    const size_t smem_need = 4 * ((size_t)k * C + (NRB - 1) * stride + k + 4) + (size_t)(C / 8) * (NRB * 16 + 16);
    if (smem_need <= 200 * 1024) {
      switch (C / 8) {
        case 32: return launch_noise_tiled<32>(reinterpret_cast<__half*>(x), src, wn, bn, B, L, Lsrc, k, stride, pad, slope, s);
        case 16: return launch_noise_tiled<16>(reinterpret_cast<__half*>(x), src, wn, bn, B, L, Lsrc, k, stride, pad, slope, s);
        case 8: return launch_noise_tiled<8>(reinterpret_cast<__half*>(x), src, wn, bn, B, L, Lsrc, k, stride, pad, slope, s);
        default: break;
      }
    }
  }
  const size_t n = (size_t)B * (C / 8) * L;
  if (dt == DT_F32)
    noise_inject_planes_kernel<float><<<blocks_for(n), 256, 0, s>>>(reinterpret_cast<float*>(x), a16, lo16, src, wn,
                                                                   bn, B, L, C, Lsrc, k, stride, pad, slope);
  else
    noise_inject_planes_kernel<__half><<<blocks_for(n), 256, 0, s>>>(reinterpret_cast<__half*>(x), a16, nullptr, src,
                                                                    wn, bn, B, L, C, Lsrc, k, stride, pad, slope);
  return cudaGetLastError();
}

// Operand planes of the long noise convs on the tensor cores (noise_convs[0]: k = 2*stride = 64 / 80 taps).
//   x[t][co] += b[co] + sum_j w[j][co] * src[t*stride + j - pad]
// is a `taps`-tap conv (taps = k / stride) over the strided frames  X[t][ci] = src[t*stride + ci - pad], ci < stride.
// The frames are written as f16 planes [B][3*SEG/8][L][8] in three segments of SEG = 64 channels: hi(src), lo(src)
// (the f16 rounding remainder) and hi(src) again -- multiplied by the weight segments (hi(w), hi(w), lo(w)) the three
// products of a two-term split, so the noise term keeps fp32-class accuracy (~2^-22) like the CUDA-core kernel it
// replaces.  One thread = one 16 B store; t is the fastest index.
__global__ void __launch_bounds__(256) source_frames_kernel(const float* __restrict__ src, __half* __restrict__ xs,
                                                            int B, int L, int Lsrc, int stride, int pad) {
  constexpr int PL = 3 * NOISE_TC_SEG / 8;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * PL * L) return;
  const int t = (int)(i % L);
  const int p = (int)((i / L) % PL);
  const int b = (int)(i / ((size_t)L * PL));
  const int seg = p / (NOISE_TC_SEG / 8), c0 = (p % (NOISE_TC_SEG / 8)) * 8;
  const float* sp = src + (size_t)b * Lsrc;
  __align__(16) __half out[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int ci = c0 + e;
    const long long n = (long long)t * stride + ci - pad;
    const float v = (ci < stride && n >= 0 && n < Lsrc) ? sp[n] : 0.f;
    const __half hi = __float2half_rn(v);
    out[e] = seg == 1 ? __float2half_rn(v - __half2float(hi)) : hi;
  }
  *reinterpret_cast<uint4*>(xs + i * 8) = *reinterpret_cast<const uint4*>(out);
}

cudaError_t launch_source_frames(const float* src, __half* xs, int B, int L, int Lsrc, int stride, int pad,
                                 cudaStream_t s) {
  if (stride < 1 || stride > NOISE_TC_SEG) return cudaErrorInvalidValue;
  const size_t n = (size_t)B * (3 * NOISE_TC_SEG / 8) * L;
  source_frames_kernel<<<blocks_for(n), 256, 0, s>>>(src, xs, B, L, Lsrc, stride, pad);
  return cudaGetLastError();
}

cudaError_t launch_conv_post_sum(const float* part, float* wave, int B, int L, int n_half, const int* tlen, int len_mul,
                                 cudaStream_t s) {
  if (n_half < 1) return cudaErrorInvalidValue;
  conv_post_sum_kernel<<<dim3((L + 255) / 256, B), 256, 0, s>>>(part, wave, B, L, n_half, tlen, len_mul);
  return cudaGetLastError();
}

cudaError_t launch_conv_post_planes(const void* x, DType dt, const float* w, float* wave, int B, int L, int C,
                                    int K, float in_slope, const int* tlen, int len_mul, cudaStream_t s) {
  if (K != 7 || C % 8 || C > 256) return cudaErrorInvalidValue;
  dim3 grid((L + CPB - 1) / CPB, B);
  const size_t smem = sizeof(float) * ((size_t)K * C + (size_t)K * (CPB + K - 1));
  if (in_slope > 1.f) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    static DeviceOnce o32, o16;
    if (dt == DT_F32) {
      if (cudaError_t e = ensure_dyn_smem(conv_post_planes_kernel<float, 7>, o32, (int)smem)) return e;
    } else {
      if (cudaError_t e = ensure_dyn_smem(conv_post_planes_kernel<__half, 7>, o16, (int)smem)) return e;
    }
  }
  if (dt == DT_F32)
    conv_post_planes_kernel<float, 7><<<grid, 256, smem, s>>>(reinterpret_cast<const float*>(x), w, wave, L, C,
                                                             in_slope, tlen, len_mul);
  else
    conv_post_planes_kernel<__half, 7><<<grid, 256, smem, s>>>(reinterpret_cast<const __half*>(x), w, wave, L, C,
                                                              in_slope, tlen, len_mul);
  return cudaGetLastError();
}

}  // namespace pg
