/*
 * polgen_rvc.h -- C ABI of the B200-native RVC synthesizer decode.
 *
 * The reference (Bebra777228/PolGen-RVC) is pure Python and has no FFI; its
 * boundary for this path is the module-object contract between the callers in
 * rvc/infer/{infer,pipeline}.py and rvc/lib/algorithm/synthesizers.Synthesizer
 * (SURVEY.md 8(b)).  Each entry point below names the reference interface it
 * stands in for.  All functions return 0 on success and a negative pg_status
 * otherwise; pg_last_error() returns a thread-local message for the last
 * failure.  Pointers marked "dev" are CUDA device pointers owned by the caller,
 * "host" are host pointers.  One handle belongs to one GPU and is not
 * thread-safe; independent handles may run concurrently.
 *
 * Tensor layouts at this boundary are time-major ("NLC"):
 *   phone   [B][T][input_dim] f32   (== the reference's (B,T,D) layout)
 *   pitch   [B][T] i64 in 1..255, f0 [B][T] f32 Hz (0 = unvoiced)
 *   lengths [B] i64, sid [B] i64
 *   eps_zp  [B][T][inter] f32  -- N(0,1) draw of synthesizers.py:174 (nullable)
 *   eps_src [B][T*upp]    f32  -- N(0,1) draw of generators.py:154   (nullable)
 *   wave    [B][T*upp]    f32  (== the reference's o[:,0,:])
 *   aux     4 x [B][T][inter] f32 : z, z_p, m_p, logs_p (nullable)
 * The Python host class hands the (B,C,T) views the reference returns back as
 * transposed views of these buffers.
 */
#ifndef POLGEN_RVC_H_
#define POLGEN_RVC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_ABI_VERSION 2
#define PG_MAX_UPS 8
#define PG_MAX_RESBLOCK_KERNELS 4
#define PG_MAX_DILATIONS 4

typedef enum pg_status {
  PG_OK = 0,
  PG_ERR_INVALID = -1,   /* bad argument / unknown tensor name / shape mismatch */
  PG_ERR_CUDA = -2,      /* CUDA runtime or driver error */
  PG_ERR_STATE = -3,     /* call order violated (e.g. infer before finalize) */
  PG_ERR_UNSUPPORTED = -4
} pg_status;

typedef enum pg_dtype { PG_F32 = 0, PG_F16 = 1, PG_I64 = 2, PG_F64 = 3 } pg_dtype;

/* Mirrors the reference constructor
 *   Synthesizer(spec_channels, segment_size, inter_channels, hidden_channels,
 *               filter_channels, n_heads, n_layers, kernel_size, p_dropout,
 *               resblock, resblock_kernel_sizes, resblock_dilation_sizes,
 *               upsample_rates, upsample_initial_channel, upsample_kernel_sizes,
 *               spk_embed_dim, gin_channels, sr, use_f0, input_dim, is_half)
 * (rvc/lib/algorithm/synthesizers.py:14-37).  Only the values the inference
 * path reads are carried; use_f0 must be 1 and resblock "1" (the non-F0
 * Generator is broken in the reference, generators.py:57). */
typedef struct pg_config {
  int32_t input_dim;                 /* 768 (v2) | 256 (v1) */
  int32_t inter_channels;            /* 192 */
  int32_t hidden_channels;           /* 192 */
  int32_t filter_channels;           /* 768 */
  int32_t n_heads;                   /* 2 */
  int32_t n_layers;                  /* 6 */
  int32_t kernel_size;               /* FFN kernel size, 3 */
  int32_t attn_window;               /* 10 (encoders.py:22) */
  int32_t gin_channels;              /* 256 */
  int32_t spk_embed_dim;             /* rows of emb_g */
  int32_t sr;                        /* 32000 | 40000 | 48000 */
  int32_t upsample_initial_channel;  /* 512 */
  int32_t n_ups;
  int32_t upsample_rates[PG_MAX_UPS];
  int32_t upsample_kernel_sizes[PG_MAX_UPS];
  int32_t n_resblock_kernels;        /* 3 */
  int32_t resblock_kernel_sizes[PG_MAX_RESBLOCK_KERNELS];
  int32_t n_dilations;               /* 3 */
  int32_t resblock_dilations[PG_MAX_RESBLOCK_KERNELS][PG_MAX_DILATIONS];
  int32_t flow_n_flows;              /* 4  (residuals.py:116) */
  int32_t flow_wn_layers;            /* 3  (synthesizers.py:104) */
  int32_t flow_wn_kernel;            /* 5 */
  int32_t flags;                     /* PG_FLAG_* */
} pg_config;

#define PG_FLAG_FORCE_SIMT 1   /* run every conv on the CUDA-core fp32 kernels (validation aid) */
#define PG_FLAG_KEEP_TAPS 2    /* keep copies of intermediates for pg_debug_fetch */
#define PG_FLAG_PROFILE 4      /* CUDA events around every conv launch (pg_profile_read) */
#define PG_FLAG_NO_PAIR_FUSION 32 /* run every ResBlock conv as its own kernel (validation twin of the fused pair kernel) */
#define PG_FLAG_LEGACY_DECODER 16 /* decoder on the time-major tcgen05 kernel (A/B comparison aid) */
#define PG_FLAG_NO_GRAPHS 64    /* launch every kernel of pg_infer directly (default: a CUDA graph per (B,T) is captured on the second call of a shape and replayed afterwards when the stream is capturable and the noise is device-drawn) */
#define PG_FLAG_NO_PLANES_SWAP 128  /* validation twin: C=128 ResBlock convs WITHOUT the swapped MMA operands (default: weights = M, 256 time rows = N where the shape qualifies) */
#define PG_FLAG_NO_PAD 256       /* do not pad T up to a launch-shape bucket (validation twin of the padded path) */
#define PG_FLAG_F32_STREAM 512   /* last decoder stage with an fp32 residual stream + f16 operand copy (validation twin of the default hi/lo f16 pair stream) */
#define PG_FLAG_NO_NOISE_FUSION 1024 /* NSF source injection as its own kernel after every upsampler (validation twin of the fused epilogue) */
#define PG_FLAG_NO_POST_FUSION 4096 /* conv_post as its own kernel over the fp32 mean (validation twin of the partial dot products the last ResBlock pair's epilogue writes) */
#define PG_FLAG_LEGACY_ATTENTION 2048 /* TextEncoder attention on the mma.sync kernel (validation twin of the tcgen05 / TMEM kernel) */
#define PG_FLAG_F16_LATENTS 8  /* TextEncoder / flow GEMMs with single-pass f16 tensor-core operands (default: fp32-accurate) */

typedef struct pg_handle_s* pg_handle;

const char* pg_last_error(void);
int pg_abi_version(void);

/* Synthesizer.__init__ (synthesizers.py:14-112) on CUDA device `device`. */
int pg_create(const pg_config* cfg, int device, pg_handle* out);

/* load_state_dict (infer.py:100): one call per tensor, reference key names with
 * weight-norm already folded by the caller into plain "<module>.weight"
 * (w = g*v/||v||).  `data` is a HOST pointer to contiguous f32. */
int pg_load_tensor(pg_handle h, const char* name, const void* data,
                   const int64_t* shape, int ndim, int dtype);

/* .eval().to(device) (infer.py:101): checks every tensor arrived, repacks the
 * weights into the kernels' layouts (tap-major fp16 K-major tiles for the
 * tcgen05 convs, polyphase form for ConvTranspose1d) and uploads them. */
int pg_finalize(pg_handle h);

/* bytes of device workspace pg_infer(B,T) will hold (grown lazily). */
size_t pg_workspace_bytes(pg_handle h, int B, int T);

/* Synthesizer.infer(phone, phone_lengths, pitch, nsff0, sid)
 * (synthesizers.py:162-188) as called from rvc/infer/pipeline.py:275.
 * All rows of the batch share T; lengths[b] <= T mask the encoder/flow exactly
 * as sequence_mask does (commons.py:89-93).  eps_* == NULL draws the noise on
 * device from Philox(seed).  Enqueued on `stream` (a cudaStream_t); returns
 * without synchronising.
 * The call is staged: inputs are copied (cudaMemcpyDefault on `stream`, so device OR pinned
 * host pointers work for phone/lengths/pitch/f0/sid/wave/aux) into handle-owned buffers whose
 * frame count is padded up to a bucket (32 / 64 / 128 frames), the launch sequence runs over
 * that padded shape with the true T as a hard end of every row (read from device memory, so
 * the result does not depend on the padding), and the waveform is copied out.  On a
 * capturable stream (not the legacy default stream) with device-drawn noise the launch
 * sequence is a CUDA graph per (B, padded T): captured on the second call of a bucket,
 * replayed afterwards for EVERY T inside it; at most pg_set_graph_cache() graphs are kept
 * (LRU).  PG_FLAG_NO_GRAPHS, explicit eps_* or the legacy stream give direct launches.
 * Row b of a batch draws its noise from Philox(seed + b*0x9E3779B97F4A7C15), indexed by
 * (t, c) only: a row's result does not depend on the rest of the batch. */
int pg_infer(pg_handle h, void* stream, int B, int T,
             const float* phone_dev, const int64_t* lengths_dev,
             const int64_t* pitch_dev, const float* f0_dev, const int64_t* sid_dev,
             const float* eps_zp_dev, const float* eps_src_dev, uint64_t seed,
             float* wave_dev, float* aux_dev);

/* The silence-split segments of a clip as ONE call (rvc/infer/pipeline.py:381-447 runs them as a
 * loop of B=1 infer calls).  Every segment is stand-alone: each layer of TextEncoder, flow AND
 * decoder treats segs[b].T as the hard end of row b (zero padding there, exactly what a B=1 call of
 * that length sees), so segs[b].wave equals pg_infer(B=1, T=segs[b].T) on that segment (same noise
 * when eps_* are given; Philox row b as documented above otherwise).  Unlike a padded pg_infer batch
 * -- which reproduces the reference's ragged-batch behaviour, where GeneratorNSF ignores the mask
 *  -- rows of different lengths do not disturb each other, and decoder tiles past a
 * row's end are skipped.  Pointers are device or pinned-host memory (copied on `stream`). */
typedef struct pg_segment {
  int32_t T;               /* frames */
  int32_t trim;            /* samples dropped at BOTH ends on copy-out (t_pad_tgt, pipeline.py:397): wave gets T*upp - 2*trim */
  int64_t sid;             /* speaker id */
  const float* phone;      /* [T][input_dim] */
  const int64_t* pitch;    /* [T] */
  const float* f0;         /* [T] */
  const float* eps_zp;     /* [T][inter]  nullable (all segments or none) */
  const float* eps_src;    /* [T*upp]     nullable */
  float* wave;             /* [T*upp - 2*trim] out */
  float* aux;              /* nullable out: z, z_p, m_p, logs_p, each [T][inter] */
} pg_segment;
int pg_infer_segments(pg_handle h, void* stream, int n, const pg_segment* segs, uint64_t seed);

/* Same call with HOST buffers (pinned recommended): copies the inputs to the
 * device, runs pg_infer, copies the waveform back and synchronises -- the
 * `.data.cpu().float().numpy()` round trip of pipeline.py:275-279. */
int pg_infer_host(pg_handle h, int B, int T,
                  const float* phone, const int64_t* lengths, const int64_t* pitch,
                  const float* f0, const int64_t* sid, uint64_t seed, float* wave);

/* Module-level entry points (the submodules a caller of the reference can
 * invoke on their own).  Same layouts/conventions as pg_infer. */

/* TextEncoder.forward (encoders.py:111-126): m_p, logs_p [B][T][inter]. */
int pg_text_encoder(pg_handle h, void* stream, int B, int T,
                    const float* phone_dev, const int64_t* lengths_dev,
                    const int64_t* pitch_dev, float* m_p_dev, float* logs_p_dev);

/* ResidualCouplingBlock.forward(reverse=True) (residuals.py:144-157):
 * z_p [B][T][inter] -> z [B][T][inter]. */
int pg_flow_reverse(pg_handle h, void* stream, int B, int T,
                    const float* z_p_dev, const int64_t* lengths_dev,
                    const int64_t* sid_dev, float* z_dev);

/* SourceModuleHnNSF.forward (nsf.py:36-40) over SineGen.forward
 * (generators.py:117-156): f0 [B][T] -> source [B][T*upp].  sine_dev
 * (nullable) receives the deterministic sine_waves*uv before noise. */
int pg_source(pg_handle h, void* stream, int B, int T, const float* f0_dev,
              const float* eps_src_dev, uint64_t seed, float* source_dev, float* sine_dev);

/* GeneratorNSF.forward (nsf.py:120-144): z [B][T][inter] (already masked),
 * source [B][T*upp], sid -> wave [B][T*upp]. */
int pg_generator(pg_handle h, void* stream, int B, int T, const float* z_dev,
                 const float* source_dev, const int64_t* sid_dev, float* wave_dev);

/* ---- the glue on either side of infer in VC.vc / VC.pipeline ---- */

/* Tail of VC.get_f0 (pipeline.py:186-201) + the casts of :379-380: f0 (Hz, already pitch-shifted;
 * f64 as the reference's numpy array, or f32 widened) -> coarse pitch 1..255 on the mel scale
 * (f0_min, f0_max = 50, 1100 in the reference) as i64, and pitchf = f0 as f32.  Bit-exact. */
int pg_coarse_pitch(pg_handle h, void* stream, int64_t n, const void* f0_dev, int f0_dtype,
                    double f0_min, double f0_max, int64_t* pitch_dev, float* pitchf_dev);

/* VC.vc's feature glue (pipeline.py:252-270): feats [n_feat_frames][input_dim] (HuBERT output after the
 * optional faiss mix) -> x2 nearest interpolate -> protect mix with feats0 (the pre-mix features;
 * applied when feats0/pitchf are given and protect < 0.5) -> phone [p_len][input_dim], where
 * p_len = min(n_audio_frames, 2*n_feat_frames) is returned in *p_len_out (pitch / pitchf are cut to
 * it by the caller, :262-263).  Bit-exact in fp32. */
int pg_prepare_features(pg_handle h, void* stream, int64_t n_feat_frames, int64_t n_audio_frames,
                        const float* feats_dev, const float* feats0_dev, const float* pitchf_dev,
                        float protect, float* phone_dev, int64_t* p_len_out);

/* Tail of VC.pipeline (pipeline.py:449-460) on the concatenated, trimmed waveform `audio` [n] (see
 * pg_segment.trim): optional AudioProcessor.change_rms (pipeline.py:31-61; librosa.feature.rms 0.10.2
 * with center=True, zero padding; linear interpolate; applied when src_audio != NULL and
 * rms_mix_rate != 1) against the 16 kHz source clip, then audio_max = max|x|/0.99, scale =
 * 32768 (/ audio_max if > 1), int16 by truncation.  audio_out (nullable, device) receives the f32 audio
 * after change_rms; pcm_out [n] may be device or pinned host memory.  The int16 step is bit-exact
 * given the f32 audio; change_rms is f32 arithmetic (relative 1e-5).  The optional librosa.resample
 * step (:454-455) is not included. */
int pg_postprocess(pg_handle h, void* stream, const float* audio_dev, int64_t n,
                   const float* src_audio_dev, int64_t n_src, int src_rate, int tgt_rate,
                   float rms_mix_rate, float* audio_out_dev, int16_t* pcm_out);

/* Copy an intermediate of the LAST pg_infer/pg_generator call to `dst_dev`
 * as f32 (parity tests tap the same points the oracle exposes: "enc.x0",
 * "enc.layer<i>", "flow.<f>", "dec.conv_pre", "dec.ups<i>", "dec.stage<i>").
 * shape_out receives up to 3 dims [B][L][C]; returns element count or <0. */
int64_t pg_debug_fetch(pg_handle h, void* stream, const char* tap, float* dst_dev,
                       int64_t capacity, int64_t* shape_out);

/* Statistics of the last call: number of kernels this library launched (for a replayed CUDA
 * graph: the kernel nodes of the graph plus the seed store). */
int64_t pg_launch_count(pg_handle h);

/* Number of (B, padded T) shapes of pg_infer currently held as instantiated CUDA graphs. */
int pg_graph_count(pg_handle h);
/* Bound on that number (default 16, env PG_GRAPH_CACHE); least recently used graphs are dropped. */
int pg_set_graph_cache(pg_handle h, int max_graphs);
/* Cap on the CTAs of the decoder's persistent conv kernels (0 = one per SM, the default; env PG_DECODER_SMS).
 * With two or more handles decoding a stream of clips on their own streams (SegmentScheduler lanes), a cap a
 * few SMs below the device's count leaves those SMs free, so the short TextEncoder / flow kernels of one
 * handle's next clip run beside another handle's decoder instead of queueing behind its persistent CTAs.
 * Results do not depend on it (the tile walk is grid-stride).  Drops the handle's captured graphs. */
int pg_set_decoder_sms(pg_handle h, int sms);
/* The padded frame count pg_infer runs a T-frame call at (the graph bucket of T). */
int pg_padded_frames(pg_handle h, int T);

/* PG_FLAG_PROFILE: device time, algorithmic FLOPs (2*B*L*Cin*Cout*K) and launch count of the
 * conv launches since the previous read, per kernel class: [0] tcgen05 channel-plane conv (decoder),
 * [1] CUDA-core conv, [2] tcgen05 time-major conv (TextEncoder / flow GEMMs, split precision).
 * Each array has 3 entries.  Synchronises on the recorded events. */
int pg_profile_read(pg_handle h, double* ms_out, double* flops_out, int64_t* launches_out);

/* PG_FLAG_PROFILE: the same records aggregated per layer shape since the previous call (the
 * records pass through pg_profile_read first).  Each row is 9 doubles: class, Cin, N (negative:
 * fused ResBlock pair), K, dilation, MT (128-row tiles per CTA tile the launch plan chose; 0 for
 * classes 1/2), launches, ms, algorithmic FLOPs.  Returns the number of rows written (<= max_rows). */
int pg_profile_table(pg_handle h, double* rows, int max_rows);

/* Single-layer entry used by the op-level parity tests and micro-benchmarks:
 * y[b][t][co] = sum_{tap,ci} w[co][ci][tap] * lrelu(x[b][t + tap*dil - pad][ci], in_slope) + bias
 * x, y are f16 [B][L][C] device buffers, w f32 host [Cout][Cin][K] (Conv1d
 * layout).  impl: 0 = CUDA-core fp32, 1 = tcgen05 time-major kernel, 2 = tcgen05
 * channel-plane kernel (the decoder's; layout converted around it).  Returns elapsed ms of
 * `iters` launches (CUDA events) in *ms_out when non-NULL. */
int pg_op_conv1d_f16(int device, int impl, int B, int L, int Cin, int Cout, int K, int dil,
                     const void* x_dev, const float* w_host, const float* bias_host,
                     float in_slope, float out_slope, const void* res_dev, void* y_dev,
                     int iters, float* ms_out);

int pg_destroy(pg_handle h);

#ifdef __cplusplus
}
#endif
#endif /* POLGEN_RVC_H_ */
