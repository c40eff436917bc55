// Harmonic source and the bandwidth-bound ends of GeneratorNSF:
//   SineGen.forward          rvc/lib/algorithm/generators.py:117-156
//   SourceModuleHnNSF.forward rvc/lib/algorithm/nsf.py:36-40
//   noise_convs[i] injection  nsf.py:93-101, :131
//   leaky_relu(0.01) -> conv_post -> tanh   nsf.py:142-143
#include <curand_kernel.h>

#include <type_traits>

#include "pg_common.cuh"

namespace pg {

// ---------------------------------------------------------------------------
// SineGen.  With harmonic_num = 0 the reference waveform is, up to rounding,
//   0.1 * sin(2*pi*frac(sum_{j<=n} rad[j // upp])) * uv + noise,
//   rad[t] = fl32(f0[t]/sr) % 1     (generators.py:126)
// (the integer cumsum_shift terms of :141-144 vanish under sin(2*pi*.);
// SURVEY.md Appendix B).  torch's CPU cumsum accumulates fp32 inputs in double,
// so the phase is accumulated in fp64 here: a frame-level exclusive prefix
//   P[t] = frac(upp * sum_{s<t} rad[s])
// (warp-shuffle scan, one block per batch row) followed by the closed form
//   phase(n) = P[t] + (r+1)*rad[t],  t = n / upp, r = n % upp.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) frame_phase_kernel(const float* __restrict__ f0,
                                                          double* __restrict__ P, int T, int upp,
                                                          float sr) {
  __shared__ double warp_tot[8];
  __shared__ double carry_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0.0;
  __syncthreads();
  for (int base = 0; base < T; base += 256) {
    const int t = base + tid;
    double v = 0.0;
    if (t < T) {
      const float rad = fmodf(__fdiv_rn(f0[(size_t)b * T + t], sr), 1.0f);
      v = (double)rad * (double)upp;
    }
    double incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    double woff = 0.0;
    for (int w = 0; w < wid; ++w) woff += warp_tot[w];
    const double carry = carry_s;
    if (t < T) {
      const double excl = carry + woff + incl - v;
      P[(size_t)b * T + t] = excl - floor(excl);
    }
    __syncthreads();
    if (tid == 255) {
      const double tot = carry + woff + incl;
      carry_s = tot - floor(tot);   // only the fractional phase matters
    }
    __syncthreads();
  }
}

__global__ void source_kernel(const float* __restrict__ f0, const double* __restrict__ P,
                              const float* __restrict__ eps, int eps_T, uint64_t seed,
                              const uint64_t* __restrict__ seeds_dev, const int* __restrict__ tlen, float lin_w,
                              float lin_b, float* __restrict__ source, float* __restrict__ sine_out, int B, int T,
                              int upp, float sr) {
  const size_t L = (size_t)T * upp;
  const size_t total = (size_t)B * L;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / L, n = i % L;
    const int t = (int)(n / upp), r = (int)(n % upp);
    if (tlen && t >= tlen[b]) {        // past the hard end of the row: the noise convs see zero padding
      source[i] = 0.f;
      if (sine_out) sine_out[i] = 0.f;
      continue;
    }
    const float f = f0[b * T + t];
    const float rad = fmodf(__fdiv_rn(f, sr), 1.0f);
    const double ph = P[b * T + t] + (double)(r + 1) * (double)rad;
    const float fr = (float)(ph - floor(ph));
    const float uv = f > 0.f ? 1.f : 0.f;
    const float sine = sinpif(2.f * fr) * 0.1f * uv;
    float e;
    if (eps) {
      e = t < eps_T ? eps[b * ((size_t)eps_T * upp) + n] : 0.f;
    } else {
      curandStatePhilox4_32_10_t st;
      curand_init((seeds_dev ? seeds_dev[b] : row_seed(seed, (int)b)) ^ 0x9E3779B97F4A7C15ull, n, 0, &st);
      e = curand_normal(&st);
    }
    const float amp = uv * 0.003f + (1.f - uv) * (0.1f / 3.f);
    const float wav = sine + amp * e;
    source[i] = tanhf(fmaf(wav, lin_w, lin_b));
    if (sine_out) sine_out[i] = sine;
  }
}

cudaError_t launch_source(const float* f0, const float* eps, int eps_T, uint64_t seed, const uint64_t* seeds_dev,
                          const int* tlen, float lin_w, float lin_b, double* frame_phase, float* source,
                          float* sine, int B, int T, int upp, int sr, cudaStream_t s) {
  frame_phase_kernel<<<B, 256, 0, s>>>(f0, frame_phase, T, upp, (float)sr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const size_t total = (size_t)B * T * upp;
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  source_kernel<<<(unsigned)blocks, 256, 0, s>>>(f0, frame_phase, eps, eps_T, seed, seeds_dev, tlen, lin_w, lin_b,
                                                 source, sine, B, T, upp, (float)sr);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// x[b][t][c] += bn[c] + sum_j wn[j][c] * src[b][t*stride + j - pad]   (nsf.py:131)
// one thread per (t, channel pair); the source window is a warp broadcast.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void noise_inject_kernel(T* __restrict__ x, const float* __restrict__ src,
                                    const float* __restrict__ wn /*[k][C]*/,
                                    const float* __restrict__ bn, int B, int L, int C, int Lsrc, int k,
                                    int stride, int pad) {
  const int c2 = C >> 1;
  const size_t total = (size_t)B * L * c2;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cp = (int)(i % c2);
  const size_t row = i / c2;
  const int t = (int)(row % L);
  const size_t b = row / L;
  const int c = cp * 2;
  float a0 = bn[c], a1 = bn[c + 1];
  const float* sp = src + b * Lsrc;
  const int s0 = t * stride - pad;
  for (int j = 0; j < k; ++j) {
    const int n = s0 + j;
    if (n < 0 || n >= Lsrc) continue;
    const float sv = sp[n];
    const float2 w = *reinterpret_cast<const float2*>(wn + (size_t)j * C + c);
    a0 = fmaf(w.x, sv, a0);
    a1 = fmaf(w.y, sv, a1);
  }
  if constexpr (sizeof(T) == 2) {
    __half2* xp = reinterpret_cast<__half2*>(x + row * C + c);
    const float2 v = __half22float2(*xp);
    *xp = __floats2half2_rn(v.x + a0, v.y + a1);
  } else {
    float2* xp = reinterpret_cast<float2*>(x + row * C + c);
    const float2 v = *xp;
    *xp = make_float2(v.x + a0, v.y + a1);
  }
}

cudaError_t launch_noise_inject(void* x, DType dt, const float* src, const float* wn, const float* bn,
                                int B, int L, int C, int Lsrc, int k, int stride, int pad,
                                cudaStream_t s) {
  const size_t total = (size_t)B * L * (C / 2);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dt == DT_F16)
    noise_inject_kernel<__half><<<grid, 256, 0, s>>>(reinterpret_cast<__half*>(x), src, wn, bn, B, L, C,
                                                     Lsrc, k, stride, pad);
  else
    noise_inject_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<float*>(x), src, wn, bn, B, L, C,
                                                    Lsrc, k, stride, pad);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// wave[b][n] = tanh(sum_{j,c} w[j][c] * lrelu(x[b][n + j - K/2][c], slope))
// 256 outputs per block; the x tile sits in shared memory with a 1-word row pad
// so the per-thread row reads are bank-conflict free.
// ---------------------------------------------------------------------------
template <int C, int K, typename T>
__global__ void __launch_bounds__(256) conv_post_kernel(const T* __restrict__ x,
                                                        const float* __restrict__ w,
                                                        float* __restrict__ wave, int L,
                                                        float in_slope) {
  constexpr int TILE = 256, W2 = C / 2, RS = W2 + 1;
  using Pair = typename std::conditional<sizeof(T) == 2, __half2, float2>::type;
  __shared__ Pair xs[(TILE + K - 1) * RS];
  __shared__ float ws[K * C];
  const int b = blockIdx.y, n0 = blockIdx.x * TILE, tid = threadIdx.x;
  for (int i = tid; i < K * C; i += 256) ws[i] = w[i];
  const Pair* xb = reinterpret_cast<const Pair*>(x + (size_t)b * L * C);
  for (int i = tid; i < (TILE + K - 1) * W2; i += 256) {
    const int r = i / W2, cw = i % W2;
    const int n = n0 + r - K / 2;
    Pair v;
    v.x = 0;
    v.y = 0;
    if (n >= 0 && n < L) v = xb[(size_t)n * W2 + cw];
    xs[r * RS + cw] = v;
  }
  __syncthreads();
  const int n = n0 + tid;
  if (n >= L) return;
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) {
#pragma unroll
    for (int cw = 0; cw < W2; ++cw) {
      float2 v;
      if constexpr (sizeof(T) == 2) v = __half22float2(xs[(tid + j) * RS + cw]);
      else v = xs[(tid + j) * RS + cw];
      v.x = v.x > 0.f ? v.x : v.x * in_slope;
      v.y = v.y > 0.f ? v.y : v.y * in_slope;
      acc = fmaf(ws[j * C + 2 * cw], v.x, acc);
      acc = fmaf(ws[j * C + 2 * cw + 1], v.y, acc);
    }
  }
  wave[(size_t)b * L + n] = tanhf(acc);
}

template <typename T>
static cudaError_t launch_conv_post_t(const T* x, const float* w, float* wave, int B, int L, int C,
                                      float in_slope, cudaStream_t s) {
  dim3 grid((L + 255) / 256, B);
  if (C == 32)
    conv_post_kernel<32, 7, T><<<grid, 256, 0, s>>>(x, w, wave, L, in_slope);
  else if (C == 16)
    conv_post_kernel<16, 7, T><<<grid, 256, 0, s>>>(x, w, wave, L, in_slope);
  else if (C == 64 && sizeof(T) == 2)
    conv_post_kernel<64, 7, __half><<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(x), w, wave, L,
                                                          in_slope);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_conv_post(const void* x, DType dt, const float* w, float* wave, int B, int L, int C,
                             int K, float in_slope, cudaStream_t s) {
  if (K != 7) return cudaErrorInvalidValue;
  if (dt == DT_F16)
    return launch_conv_post_t(reinterpret_cast<const __half*>(x), w, wave, B, L, C, in_slope, s);
  return launch_conv_post_t(reinterpret_cast<const float*>(x), w, wave, B, L, C, in_slope, s);
}

}  // namespace pg
