// The host-numpy glue on either side of Synthesizer.infer in the reference's VC pipeline
// (SURVEY.md 8(f) ranks 2 and 3), as bandwidth kernels so features, pitch and audio stay on the GPU:
//   coarse pitch quantisation          rvc/infer/pipeline.py:186-201 (get_f0 tail) + :379-380
//   x2 nearest interpolate + protect   rvc/infer/pipeline.py:252-270 (VC.vc)
//   change_rms                         rvc/infer/pipeline.py:31-61  (librosa.feature.rms 0.10.2 + F.interpolate)
//   peak normalise + int16             rvc/infer/pipeline.py:456-460
// Arithmetic mirrors the reference's numpy 1.23.5 / torch expressions operation by operation (no fused
// multiply-adds, same rounding points), so the integer outputs are bit-exact.
#include <math.h>

#include "pg_common.cuh"

namespace pg {

namespace {

// f0_mel = 1127*ln(1 + f0/700); >0: (f0_mel - min)*254/(max - min) + 1; <=1 -> 1; >255 -> 255; rint.
// The reference's f0 is a float64 numpy array (RMVPE.decode / np.interp), so everything is double; an f32
// input is widened first (what numpy does with a float64 array holding the same values).
template <typename T>
__global__ void coarse_pitch_kernel(const T* __restrict__ f0, int64_t n, double mel_min, double mel_span,
                                    int64_t* pitch, float* pitchf) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double f = (double)f0[i];
  double mel = 1127.0 * log(1.0 + f / 700.0);
  if (mel > 0.0) mel = (mel - mel_min) * 254.0 / mel_span + 1.0;
  if (mel <= 1.0) mel = 1.0;
  if (mel > 255.0) mel = 255.0;
  pitch[i] = (int64_t)rint(mel);             // np.rint: half to even
  pitchf[i] = (float)f;                      // torch.tensor(pitchf).float()
}

// feats (1, Th, D) -> F.interpolate(scale_factor=2) (nearest: out[t] = in[t // 2]) -> protect mix:
//   pitchff = 1 where pitchf > 0, then `protect` where pitchf < 1;  out = feats*pitchff + feats0*(1 - pitchff)
__global__ void prepare_features_kernel(const float* __restrict__ feats, const float* __restrict__ feats0,
                                        const float* __restrict__ pitchf, float protect, int do_protect,
                                        float* __restrict__ out, int64_t p_len, int D) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int D4 = D / 4;
  if (i >= p_len * D4) return;
  const int64_t t = i / D4;
  const int c4 = (int)(i - t * D4);
  const float4 a = reinterpret_cast<const float4*>(feats + (t >> 1) * D)[c4];
  float4 o = a;
  if (do_protect) {
    const float pf = pitchf[t];
    float w = pf > 0.f ? 1.f : pf;      // pitchff[pitchf > 0] = 1
    if (pf < 1.f) w = protect;          // pitchff[pitchf < 1] = protect
    const float u = __fsub_rn(1.f, w);
    const float4 b = reinterpret_cast<const float4*>(feats0 + (t >> 1) * D)[c4];
    o.x = __fadd_rn(__fmul_rn(a.x, w), __fmul_rn(b.x, u));
    o.y = __fadd_rn(__fmul_rn(a.y, w), __fmul_rn(b.y, u));
    o.z = __fadd_rn(__fmul_rn(a.z, w), __fmul_rn(b.z, u));
    o.w = __fadd_rn(__fmul_rn(a.w, w), __fmul_rn(b.w, u));
  }
  reinterpret_cast<float4*>(out + t * D)[c4] = o;
}

// librosa.feature.rms(y, frame_length, hop_length, center=True, pad_mode="constant"): frame j covers
// samples [j*hop - frame/2, j*hop + frame/2) of y (zeros outside), rms = sqrt(mean(x^2)) in f32.
__global__ void __launch_bounds__(256) frame_rms_kernel(const float* __restrict__ y, int64_t n, int frame, int hop,
                                                        float* __restrict__ rms, int n_frames) {
  __shared__ double part[8];
  const int j = blockIdx.x;
  const int64_t lo = (int64_t)j * hop - frame / 2;
  double s = 0.0;       // exact-ish accumulation: within 1 ulp of any f32 summation order
  for (int i = threadIdx.x; i < frame; i += blockDim.x) {
    const int64_t k = lo + i;
    if (k >= 0 && k < n) {
      const float v = y[k];
      s += (double)__fmul_rn(v, v);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += part[w];
    rms[j] = sqrtf((float)(t / frame));
  }
}

// F.interpolate(size=N, mode="linear", align_corners=False) of a length-n array at position i
__device__ __forceinline__ float lerp_at(const float* __restrict__ a, int n, float scale, int64_t i) {
  float src = scale * ((float)i + 0.5f) - 0.5f;       // area_pixel_compute_source_index
  if (src < 0.f) src = 0.f;
  const int i0 = (int)src;
  const int i1 = i0 + (i0 < n - 1 ? 1 : 0);
  const float l1 = src - (float)i0, l0 = 1.f - l1;
  return __fadd_rn(__fmul_rn(l0, a[i0]), __fmul_rn(l1, a[i1]));
}

// out = x * (rms1^(1-rate) * max(rms2, 1e-6)^(rate-1)); also the running max |out| (as float bits)
__global__ void change_rms_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ rms1, int n1,
                                  const float* __restrict__ rms2, int n2, float rate, float* __restrict__ out,
                                  unsigned int* __restrict__ max_bits) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float v = 0.f;
  if (i < n) {
    const float r1 = lerp_at(rms1, n1, (float)n1 / (float)n, i);
    const float r2 = fmaxf(lerp_at(rms2, n2, (float)n2 / (float)n, i), 1e-6f);
    const float g = __fmul_rn(powf(r1, 1.f - rate), powf(r2, rate - 1.f));
    v = __fmul_rn(x[i], g);
    out[i] = v;
  }
  float m = fabsf(v);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(max_bits, __float_as_uint(m));   // non-negative floats order as uints
}

__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, unsigned int* __restrict__ max_bits) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(max_bits, __float_as_uint(m));
}

// audio_max = max|x| / 0.99; max_int16 = 32768 (/ audio_max if audio_max > 1); (x * max_int16).astype(int16)
__global__ void to_int16_kernel(const float* __restrict__ x, int64_t n, const unsigned int* __restrict__ max_bits,
                                int16_t* __restrict__ pcm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float audio_max = __fdiv_rn(__uint_as_float(*max_bits), 0.99f);
  float scale = 32768.f;
  if (audio_max > 1.f) scale = __fdiv_rn(32768.f, audio_max);
  pcm[i] = (int16_t)(int)__fmul_rn(x[i], scale);     // C cast: truncation toward zero, as numpy's astype
}

unsigned blocks_for(int64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

cudaError_t launch_coarse_pitch(const void* f0, int is_f64, int64_t n, double f0_min, double f0_max, int64_t* pitch,
                                float* pitchf, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  // f0_mel_min / f0_mel_max are np.float64 scalars (pipeline.py:150-151); the span is taken in double
  const double mel_min = 1127.0 * log(1.0 + f0_min / 700.0), mel_max = 1127.0 * log(1.0 + f0_max / 700.0);
  if (is_f64)
    coarse_pitch_kernel<double><<<blocks_for(n), 256, 0, s>>>(reinterpret_cast<const double*>(f0), n, mel_min,
                                                              mel_max - mel_min, pitch, pitchf);
  else
    coarse_pitch_kernel<float><<<blocks_for(n), 256, 0, s>>>(reinterpret_cast<const float*>(f0), n, mel_min,
                                                             mel_max - mel_min, pitch, pitchf);
  return cudaGetLastError();
}

cudaError_t launch_prepare_features(const float* feats, const float* feats0, const float* pitchf, float protect,
                                    float* out, int64_t p_len, int D, cudaStream_t s) {
  if (D % 4) return cudaErrorInvalidValue;
  if (p_len <= 0) return cudaSuccess;
  const int do_protect = feats0 != nullptr && pitchf != nullptr && protect < 0.5f;
  prepare_features_kernel<<<blocks_for(p_len * (D / 4)), 256, 0, s>>>(feats, feats0, pitchf, protect, do_protect,
                                                                     out, p_len, D);
  return cudaGetLastError();
}

cudaError_t launch_frame_rms(const float* y, int64_t n, int rate, float* rms, int n_frames, cudaStream_t s) {
  frame_rms_kernel<<<n_frames, 256, 0, s>>>(y, n, rate / 2 * 2, rate / 2, rms, n_frames);
  return cudaGetLastError();
}

cudaError_t launch_change_rms(const float* x, int64_t n, const float* rms1, int n1, const float* rms2, int n2,
                              float rate, float* out, unsigned int* max_bits, cudaStream_t s) {
  change_rms_kernel<<<blocks_for(n), 256, 0, s>>>(x, n, rms1, n1, rms2, n2, rate, out, max_bits);
  return cudaGetLastError();
}

cudaError_t launch_absmax(const float* x, int64_t n, unsigned int* max_bits, cudaStream_t s) {
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  absmax_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, n, max_bits);
  return cudaGetLastError();
}

cudaError_t launch_to_int16(const float* x, int64_t n, const unsigned int* max_bits, int16_t* pcm, cudaStream_t s) {
  to_int16_kernel<<<blocks_for(n), 256, 0, s>>>(x, n, max_bits, pcm);
  return cudaGetLastError();
}

}  // namespace pg
