// Shared declarations of the polgen-rvc_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

namespace pg {

// cudaFuncSetAttribute and the SM count are PER DEVICE: a process may own handles on several GPUs
// (Synthesizer.to('cuda:1') after an engine on cuda:0), so one-time launcher setup is remembered per
// (kernel instantiation, device).  Thread-safe: handles on different devices may run concurrently.
constexpr int PG_MAX_DEVICES = 64;
struct DeviceOnce {
  std::atomic<int> bytes[PG_MAX_DEVICES] = {};
};
// opt in to `bytes` of dynamic shared memory for `kernel` on the current device (idempotent)
template <class Kernel>
inline cudaError_t ensure_dyn_smem(Kernel kernel, DeviceOnce& once, int bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const bool tracked = dev >= 0 && dev < PG_MAX_DEVICES;
  if (tracked && once.bytes[dev].load(std::memory_order_acquire) >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  if (tracked) {
    int cur = once.bytes[dev].load(std::memory_order_relaxed);
    while (cur < bytes && !once.bytes[dev].compare_exchange_weak(cur, bytes, std::memory_order_release)) {}
  }
  return cudaSuccess;
}
// SM count of the CURRENT device (cached per device)
int device_sm_count();

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------
// A decode is ~185 short launches in one stream.  The tensor-core kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: every CTA executes griddepcontrol.launch_dependents first thing,
// so the NEXT kernel of the stream is launched while this one runs; its CTAs get an SM as soon as one of ours exits,
// run their prologue (barrier init, TMEM allocation, tensor-map prefetch, bias staging, weight TMA -- constants only)
// and block in griddepcontrol.wait until this grid has completed and flushed.  Only the roles that touch activations
// in global memory wait; the weight producer and the MMA issuer never do.  Launch latency and prologue of every
// kernel thus overlap the tail of its predecessor.  Kernels launched the plain way remain full barriers, and
// griddepcontrol.wait is a no-op in a kernel launched without the attribute.  PG_PDL=0 disables the attribute.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// thread-local error string behind pg_last_error()
void set_error(const std::string& msg);

#define PG_CUDA_CHECK(expr)                                                         \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::pg::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));          \
      return PG_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

// ACT_GATE (tcgen05 time-major kernel, f32 output only): output columns come in (tanh, sigmoid) pairs --
//   y[:, j] = tanh(v[2j]) * sigmoid(v[2j+1]), Cout / 2 columns are written (commons.py:79-86 fused into the in-layer GEMM)
enum Act : int { ACT_NONE = 0, ACT_LRELU = 1, ACT_RELU = 2, ACT_TANH = 3, ACT_GATE = 4 };

// One conv / GEMM layer over time-major activations.
//   in row (b, t + tap*dil - pad), channel x_coff + ci, row pitch x_ld elements
//   y = acc_old*accumulate + out_scale * (sum + bias[co] + bbias[b][co] + res_scale*res)
//   then activation, then row mask (t < lens[b]).
struct ConvArgs {
  const void* x = nullptr; int x_ld = 0; int x_coff = 0;
  int B = 0, L_in = 0, L_out = 0;
  int Cin = 0, Cout = 0, K = 1, dil = 1, pad = 0;
  const float* w = nullptr;          // SIMT: [K][Cin][Cout] f32
  const void* w16 = nullptr;         // UMMA: [K][Cout][Cin] f16 (K-major tiles)
  const void* w16s = nullptr;        // UMMA split mode: [2][K][Cout][Cin] f16, hi then lo (w = hi + lo)
  int split = 0;                     // two-term f16 operands, 3 MMAs per product (fp32-class accuracy)
  int out_f32 = 0;                   // set by the launcher: f32 outputs need the wider epilogue staging
  const float* bias = nullptr;       // [Cout]
  const float* bbias = nullptr; int bbias_ld = 0;   // [B][bbias_ld]
  float in_slope = 1.f; int in_mask = 0;
  const int* lens = nullptr;         // [B] valid rows (frames) for in_mask/out_mask
  int act = ACT_NONE; float out_slope = 1.f; int out_mask = 0;
  const void* res = nullptr; int res_ld = 0; int res_coff = 0; float res_scale = 1.f;
  void* y = nullptr; int y_ld = 0; int y_coff = 0;
  int accumulate = 0; float out_scale = 1.f;
  // tcgen05 path only: per Cout-tile bitmask of taps with non-zero weights (polyphase upsampler)
  const uint32_t* tapmask = nullptr;
};

enum DType : int { DT_F32 = 0, DT_F16 = 1 };

// ---- launchers (each returns cudaGetLastError()) --------------------------
cudaError_t launch_conv_simt(const ConvArgs& a, DType in_dt, DType out_dt, cudaStream_t s);

// tcgen05 implicit-GEMM conv over f16 activations (pg_conv_umma.cu).
bool umma_conv_supported(const ConvArgs& a);
int umma_pick_nt(int cout);   // Cout tile the tcgen05 conv will use (0 = unsupported)
cudaError_t launch_conv_umma(const ConvArgs& a, DType in_dt, DType out_dt, cudaStream_t s);

// ---- channel-plane layout (decoder): f16/f32 [B][C/8][L][8] ------------------------------
// One conv layer over plane activations on the tcgen05 kernel (pg_conv_planes.cu):
//   v = out_scale * (sum_{tap,ci} W[tap][n][ci] * x[b][t + tap*dil - pad][ci] + bias[n] + bbias[b][n]
//                    + raw(res16) + res32) + accin
//   out32 = v ; out16 = lrelu(v, out16_slope)           (either may be null)
// Column n of the GEMM maps to output row t*row_mul + n / Cout_real, channel n % Cout_real
// (row_mul > 1: polyphase ConvTranspose1d).  res16 is stored post-activation: raw = v<0 ? v*res_inv : v.
struct PlaneConvArgs {
  const __half* x = nullptr; int B = 0, L = 0, Cin = 0;
  const int* lens = nullptr; int in_mask = 0;            // input rows >= lens[b] read as zero
  const int* tlen = nullptr; int len_mul = 1;            // hard end of row b: input rows >= tlen[b]*len_mul read as
                                                          // zero and tiles that start there are skipped
  const void* w16 = nullptr;                              // [K][N][Cin] f16
  int N = 0, K = 1, dil = 1, pad = 0;
  const float* bias = nullptr; const float* bbias = nullptr; int bbias_ld = 0;
  const uint32_t* tapmask = nullptr;
  int Cout_real = 0, row_mul = 1;
  const __half* res16 = nullptr; float res_inv = 1.f; const float* res32 = nullptr;
  const __half* accin16 = nullptr; const float* accin32 = nullptr;
  __half* out16 = nullptr; float out16_slope = 1.f; float* out32 = nullptr;
  float out_scale = 1.f;
  // NSF source injection fused into the epilogue (nsf.py:131; the polyphase upsamplers): adds
  //   nz_b[co] + sum_j nz_w[j][co] * nz_src[b][t_out*nz_stride + j - nz_pad]   (samples outside [0, nz_len) are 0)
  const float* nz_src = nullptr; const float* nz_w = nullptr; const float* nz_b = nullptr;
  int nz_k = 0, nz_stride = 1, nz_pad = 0, nz_len = 0;
  __half* out_lo = nullptr;   // with out16: the value leaves as the hi/lo f16 pair (last decoder stage)
  int grid_cap = 0;   // > 0: at most this many persistent CTAs (pg_set_decoder_sms: SMs left free for another lane's small kernels)
  int swap = 1;   // operand-swapped MMA where the shape qualifies (C = 128, MT = 2); 0: PG_FLAG_NO_PLANES_SWAP twin
};
bool plane_conv_supported(const PlaneConvArgs& a);
int plane_conv_mt(const PlaneConvArgs& a);   // 128-row tiles per CTA tile the launch plan picks (0: unsupported)
int plane_pick_nt(int n);
cudaError_t launch_conv_planes(const PlaneConvArgs& a, cudaStream_t s);

// Fused ResBlock pair (pg_pair_planes.cu), C in {32, 64}:
//   tmp = lrelu(conv_dil(x; w1) + bias1, 0.1) (zero outside [0, L));  v = conv_1(tmp; w2) + bias2 + residual
//   residual = raw(x) (x is L-form: min(x, x*res_inv)) or the fp32 stream res32
//   v = out_scale * v + accin ; out32 = v ; out16 = lrelu(v, out16_slope)
struct PairConvArgs {
  const __half* x = nullptr; int B = 0, L = 0, C = 0, K = 1, dil = 1;
  const int* tlen = nullptr; int len_mul = 1;              // hard end of row b at tlen[b]*len_mul (zero beyond)
  const void* w1 = nullptr; const void* w2 = nullptr;      // [K][C][C] f16 each
  const float* bias1 = nullptr; const float* bias2 = nullptr;
  const float* res32 = nullptr; float res_inv = 1.f;
  // hi/lo stream (last decoder stage): x is the hi plane, x_lo its f16 remainder; the residual is raw(x + x_lo)
  // and, with out_lo, the output is written as the same two planes (out16 = hi, out_lo = lo)
  const __half* x_lo = nullptr; __half* out_lo = nullptr;
  const __half* accin16 = nullptr; const float* accin32 = nullptr;
  __half* out16 = nullptr; float out16_slope = 1.f; float* out32 = nullptr;
  float out_scale = 1.f;
  int grid_cap = 0;   // see PlaneConvArgs::grid_cap
  // conv_post fused into the LAST pair of the last stage (hi/lo stream, out32 form): instead of storing the fp32 mean,
  // E2 writes per row the K_POST partial dot products  part[h][j][b][t] = sum_{c in half h} post_w[j][c] * lrelu(v[c], post_slope)
  // of its 16-channel item (h = channel half); launch_conv_post_sum adds the shifted partials and applies tanh
  const float* post_w = nullptr; float* post_part = nullptr; float post_slope = 1.f;
};
constexpr int K_POST = 7;
// wave[b][t] = tanh(sum_h sum_j part[h][j][b][t + j - K_POST/2]) over rows inside [0, hard end); n_half = C / 16
cudaError_t launch_conv_post_sum(const float* part, float* wave, int B, int L, int n_half, const int* tlen, int len_mul,
                                 cudaStream_t s);
bool pair_conv_supported(const PairConvArgs& a);
int pair_conv_mt(const PairConvArgs& a);
cudaError_t launch_pair_planes(const PairConvArgs& a, cudaStream_t s);

// time-major [B][L][x_ld] (channels x_coff..x_coff+C) -> planes f16, rows >= lens[b] zeroed when lens,
// stored as lrelu(x, slope)
cudaError_t launch_nlc_to_planes(const void* x, DType dt, int x_ld, int x_coff, __half* y, int B, int L,
                                 int C, const int* lens, float slope, cudaStream_t s);
// planes (f16 or f32) -> time-major f32 [B][L][C]; inv_slope undoes an L-form storage (1 = raw)
cudaError_t launch_planes_to_nlc(const void* x, DType dt, float* y, int B, int L, int C, float inv_slope,
                                 cudaStream_t s);
// planes f16 -> time-major f16 [B][L][C]
cudaError_t launch_planes_to_nlc_f16(const __half* x, __half* y, int B, int L, int C, cudaStream_t s);
// x_raw (planes, f16 or f32, in place) += bn[c] + sum_j wn[j][c] * src[b][t*stride + j - pad];
// a16 = lrelu(x, slope) (f16 planes); f32 raw kept in x when dt == DT_F32 -- unless lo16 is given: then the
// L-form value leaves as the hi/lo pair (a16, lo16) and x is not written back
cudaError_t launch_noise_inject_planes(void* x, DType dt, __half* a16, __half* lo16, const float* src,
                                       const float* wn, const float* bn, int B, int L, int C, int Lsrc, int k,
                                       int stride, int pad, float slope, cudaStream_t s);
// strided source frames as split f16 operand planes [B][3*NOISE_TC_SEG/8][L][8] (hi | lo | hi segments) for the
// tensor-core form of a long noise conv (pg_planes.cu)
constexpr int NOISE_TC_SEG = 64;
cudaError_t launch_source_frames(const float* src, __half* xs, int B, int L, int Lsrc, int stride, int pad,
                                 cudaStream_t s);
// wave = tanh(conv_post(lrelu(x, in_slope))) over planes (f16 or f32)
// tlen (nullable): rows >= tlen[b]*len_mul read as zero (hard end of row b)
cudaError_t launch_conv_post_planes(const void* x, DType dt, const float* w /*[K][C]*/, float* wave, int B,
                                    int L, int C, int K, float in_slope, const int* tlen, int len_mul,
                                    cudaStream_t s);

// Per-call scalars the kernels of a (possibly replayed) launch sequence read from DEVICE memory, so a
// CUDA graph captured for a padded (B, Tp) bucket serves every true length inside the bucket:
//   meta[0] = T_true (frames the caller passed), meta[1] = ragged (rows are stand-alone segments)
//   seeds[b] = Philox seed of batch row b (row_seed(seed, b))
struct CallMeta { int t_true; int ragged; };
__host__ __device__ inline uint64_t row_seed(uint64_t seed, int b) {
  return seed + (uint64_t)b * 0x9E3779B97F4A7C15ull;
}
cudaError_t launch_set_call(CallMeta* meta, uint64_t* seeds, int B, int t_true, int ragged, uint64_t seed,
                            cudaStream_t s);
// lengths / speaker ids of up to N rows passed BY VALUE (kernel parameters): written to the int64 arrays
struct RowScalars { static constexpr int N = 128; int len[N]; int sid[N]; };
cudaError_t launch_set_rows(int64_t* lengths, int64_t* sid, const RowScalars& rs, int n, cudaStream_t s);
// lens32[b] = clamp(lengths[b], 0, T_true): the sequence_mask of the TextEncoder / flow (commons.py:89-93);
// tlen32[b] = hard end of row b for the decoder (every layer zero-pads there): lens32[b] when ragged,
// else T_true.  T is the row pitch of `pitch` (the padded frame count); meta == nullptr: T_true = T, dense.
cudaError_t launch_prepare_ints(const int64_t* lengths, const int64_t* pitch, const int64_t* sid,
                                const CallMeta* meta, int* lens32, int* tlen32, int* pitch32, int* sid32,
                                int B, int T, int n_spk, cudaStream_t s);
// y[b][n] = bias[n] + sum_k w[n][k] * emb[sid[b]][k]
cudaError_t launch_cond_gemv(const float* emb, const int* sid, const float* w, const float* bias,
                             float* y, int B, int Kdim, int N, cudaStream_t s);
// x = lrelu((x + emb_pitch[pitch]) * scale, 0.1) * mask
cudaError_t launch_embed_finish(float* x, const float* emb_pitch, const int* pitch, const int* lens,
                                int B, int T, int H, float scale, cudaStream_t s);
// x = LayerNorm_c(x + y) * (final_mask ? mask : 1)
cudaError_t launch_add_layernorm(float* x, const float* y, const float* gamma, const float* beta,
                                 int rows, int H, cudaStream_t s);
// relative-position windowed attention; qkv [B][T][3H] -> out [B][T][H]
cudaError_t launch_rel_attention(const float* qkv, const float* rel_k, const float* rel_v,
                                 const int* lens, float* out, int B, int T, int H, int n_heads,
                                 int window, cudaStream_t s);
// same on tensor cores (pg_attention.cu) with two-term f16 operand splitting (fp32-class accuracy): tcgen05 / TMEM
// kernel (legacy = 0, the product path) or its mma.sync validation twin (legacy = 1, PG_FLAG_LEGACY_ATTENTION);
// scratch holds the split, pre-tiled Q / K / V copies and the split-key partials (rel_attention_scratch_bytes)
size_t rel_attention_scratch_bytes(int B, int T, int H);
cudaError_t launch_rel_attention_mma(const float* qkv, const float* rel_k, const float* rel_v,
                                     const int* lens, float* out, void* scratch, int B, int T, int H,
                                     int n_heads, int window, int legacy, cudaStream_t s);
// z_p = (m + exp(logs) * eps * 0.66666) * mask ; stats [B][T][2C] -> m, logs, z_p, z(copy)
// eps (nullable) is [B][eps_T][C] (rows >= eps_T draw 0); otherwise Philox(seeds[b] | row_seed(seed, b)),
// subsequence t*C + c: a row's noise does not depend on the batch around it or on the padded T
cudaError_t launch_reparam(const float* stats, const float* eps, int eps_T, uint64_t seed,
                           const uint64_t* seeds_dev, const int* lens, float* m_p, float* logs_p, float* z_p,
                           float* z, int B, int T, int C, cudaStream_t s);
// acts = tanh(a[:, :H]) * sigmoid(a[:, H:])
cudaError_t launch_gate(const float* a, float* acts, int64_t rows, int H, cudaStream_t s);

// harmonic source (pg_source.cu)
// eps (nullable) is [B][eps_T*upp]; tlen (nullable): samples >= tlen[b]*upp are written as 0 (hard end of
// the row: the noise convs then see the zero padding a stand-alone run of that length would)
cudaError_t launch_source(const float* f0, const float* eps, int eps_T, uint64_t seed, const uint64_t* seeds_dev,
                          const int* tlen, float lin_w, float lin_b, double* frame_phase, float* source,
                          float* sine, int B, int T, int upp, int sr, cudaStream_t s);
// x[b][t][c] += bn[c] + sum_j wn[c][j] * src[b][t*stride + j - pad]
cudaError_t launch_noise_inject(void* x, DType dt, const float* src, const float* wn, const float* bn,
                                int B, int L, int C, int Lsrc, int k, int stride, int pad,
                                cudaStream_t s);
// wave = tanh(conv_post(lrelu(x, 0.01)))
cudaError_t launch_conv_post(const void* x, DType dt, const float* w /*[K][C]*/, float* wave, int B,
                             int L, int C, int K, float in_slope, cudaStream_t s);
cudaError_t launch_cast_f16_to_f32(const __half* x, float* y, int64_t n, cudaStream_t s);

// ---- pipeline glue around Synthesizer.infer (pg_pipeline.cu; rvc/infer/pipeline.py) ----
cudaError_t launch_coarse_pitch(const void* f0, int is_f64, int64_t n, double f0_min, double f0_max, int64_t* pitch,
                                float* pitchf, cudaStream_t s);
cudaError_t launch_prepare_features(const float* feats, const float* feats0, const float* pitchf, float protect,
                                    float* out, int64_t p_len, int D, cudaStream_t s);
cudaError_t launch_frame_rms(const float* y, int64_t n, int rate, float* rms, int n_frames, cudaStream_t s);
cudaError_t launch_change_rms(const float* x, int64_t n, const float* rms1, int n1, const float* rms2, int n2,
                              float rate, float* out, unsigned int* max_bits, cudaStream_t s);
cudaError_t launch_absmax(const float* x, int64_t n, unsigned int* max_bits, cudaStream_t s);
cudaError_t launch_to_int16(const float* x, int64_t n, const unsigned int* max_bits, int16_t* pcm, cudaStream_t s);

}  // namespace pg
