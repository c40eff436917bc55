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

// nsf.py:142-143  wave = tanh(conv_post(leaky_relu(x)))  (Cout = 1, no bias).  One block = 256 output samples: the
// (256 + K - 1)-row window of every plane is staged RAW in shared memory by 16-byte cp.async copies (all in flight at
// once, no register staging; rows outside [0, hard row end) are zero-filled by the copy itself), then one thread per
// sample walks the K x C taps from shared memory, applying the leaky-ReLU on the way.  (The round-1 kernel re-read
// every row K times through L1.)
constexpr int CPB = 256;
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int n = valid ? 16 : 0;       // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
template <typename T, int K>
__global__ void __launch_bounds__(CPB) conv_post_planes_kernel(const T* __restrict__ x,
                                                               const float* __restrict__ w /*[K][C]*/,
                                                               float* __restrict__ wave, int L, int C,
                                                               float in_slope, const int* __restrict__ tlen,
                                                               int len_mul) {
  extern __shared__ __align__(16) float cps[];
  constexpr int RB = 8 * (int)sizeof(T);          // bytes per (row, plane): 32 (f32) / 16 (f16)
  constexpr int PIECES = RB / 16;
  const int CP = C / 8, ROWS = CPB + K - 1;
  float* ws = cps;                                // [K][C]
  uint8_t* xs = reinterpret_cast<uint8_t*>(cps + K * C);   // [CP][ROWS] x RB bytes, raw
  const int b = blockIdx.y, t0 = blockIdx.x * CPB;
  const int t_hi = tlen ? min(L, tlen[b] * len_mul) : L;   // hard end of the row
  const uint8_t* xb = reinterpret_cast<const uint8_t*>(x) + (size_t)b * CP * (size_t)L * RB;
  for (int i = threadIdx.x; i < CP * ROWS * PIECES; i += CPB) {
    const int piece = i % PIECES, rr = i / PIECES;
    const int pl = rr / ROWS, r = rr - pl * ROWS;
    const int n = t0 + r - K / 2;
    const bool valid = n >= 0 && n < t_hi;
    cp_async16_zfill(xs + (size_t)rr * RB + piece * 16, xb + ((size_t)pl * L + (valid ? n : 0)) * RB + piece * 16, valid);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = threadIdx.x; i < K * C; i += CPB) ws[i] = w[i];
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= L) return;
  float acc = 0.f;
  for (int pl = 0; pl < CP; ++pl) {
    const T* row = reinterpret_cast<const T*>(xs + ((size_t)pl * ROWS + threadIdx.x) * RB);
#pragma unroll
    for (int j = 0; j < K; ++j) {
      float v[8];
      load8(row + j * 8, v);
      const float* wj = ws + j * C + pl * 8;
#pragma unroll
      for (int q = 0; q < 8; ++q) acc = fmaf(wj[q], fmaxf(v[q], v[q] * in_slope), acc);   // slope <= 1
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

cudaError_t launch_conv_post_planes(const void* x, DType dt, const float* w, float* wave, int B, int L, int C,
                                    int K, float in_slope, const int* tlen, int len_mul, cudaStream_t s) {
  if (K != 7 || C % 8 || C > 256) return cudaErrorInvalidValue;
  dim3 grid((L + CPB - 1) / CPB, B);
  const size_t smem = sizeof(float) * (size_t)K * C + (size_t)(C / 8) * (CPB + K - 1) * (dt == DT_F32 ? 32 : 16);
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
