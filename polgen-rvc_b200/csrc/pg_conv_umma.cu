// tcgen05 / TMEM implicit-GEMM Conv1d over time-major activations (sm_100a).
//
// One kernel serves every dense conv / GEMM of the path: the ResBlock convs
// (reference rvc/lib/algorithm/residuals.py:45-53), the polyphase form of the
// ConvTranspose1d upsamplers (nsf.py:80-91), conv_pre (nsf.py:64-66), and the
// 1x1 / k3 / k5 convs of the TextEncoder and the flow (encoders.py, attentions.py,
// modules.py):
//   y[b][t][co] = epi( sum_{tap,ci} W[tap][co][ci] * act_in(x[b][t + tap*dil - pad][ci]) )
//
// Mapping: D[M = 128 time rows, N = Cout tile] += A * B per tap and
// per 16-channel K slice, f16 operands, fp32 accumulation in TMEM.
//   A = activations.  The CTA stages the window of MT*128 + (K-1)*dil rows in
//       shared memory as channel-chunk planes [Cin/8][rows][8 ch] -- the UMMA
//       no-swizzle K-major canonical layout with SBO = 128 B -- converting to
//       f16 and applying the input mask / pre-activation leaky-ReLU on the way
//       in.  Rows are 16 B apart inside a plane, so a tap/dilation shift is just
//       `start_address += shift*16`: zero-copy implicit im2col, every tap and
//       every one of the MT row tiles re-reads the same window.  Wide inputs
//       (Cin > 256) go through a double-buffered ring of 256-channel groups.
//   B = weights, [tap][Cout][Cin] f16 in global, streamed by TMA (128B swizzle,
//       64-channel chunks; 64B swizzle, 32-channel chunks when Cin % 64 != 0)
//       through an mbarrier ring; each stage feeds MT accumulators.
// Roles: warps 0-3 stage A then run the epilogue (TMEM -> regs -> bias /
// per-batch bias / residual / scale / accumulate / activation / mask -> global),
// warp 4 is the TMA producer, warp 5 owns TMEM and issues the MMAs.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>

#include "pg_common.cuh"

namespace pg {

namespace {

constexpr int BM = 128;              // rows per accumulator tile
constexpr int NTHREADS = 192;
constexpr int GROUP_PLANES = 32;     // 256 channels per A group

// ------------------------------ PTX wrappers -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (sticky error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("pg_conv_umma: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x,
             blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (f16 operands, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------ descriptors --------------------------------
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//  [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
constexpr uint32_t LAYOUT_NONE = 0, LAYOUT_SW128 = 2, LAYOUT_SW64 = 4;

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 @4, a/b
// format F16 (0) @7/@10, K-major both, N>>3 @17, M>>4 @24.
inline uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct UmmaParams {
  const void* x; int x_ld, x_coff;
  const void* res; int res_ld, res_coff; float res_scale;
  void* y; int y_ld, y_coff;
  const float* bias; const float* bbias; int bbias_ld;
  const int* lens; int in_mask, out_mask;
  const uint32_t* tapmask;     // [n_tiles] bit t set = tap t has non-zero weights (nullable)
  int B, L;
  int Cin, Cout, K, dil, pad;
  int NT;                       // Cout tile (accumulator columns)
  float in_slope, out_slope, out_scale;
  int act, accumulate;
  int plane_bytes;              // A plane pitch (16 B * (rows_alloc + 1))
  int KC;                       // channels per weight stage: 64 (SW128) or 32 (SW64)
  int PG, n_groups, a_bufs;     // planes per A group, groups, group buffers
  int stages, stage_bytes;      // weight ring
  int tmem_cols;
  uint32_t idesc;
};

template <typename T> struct VecIO;
template <> struct VecIO<__half> {
  // 8 channels -> 8 floats
  static __device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store8(__half* p, const float (&v)[8]) {
    uint4 q;
    __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = q;
  }
};
template <> struct VecIO<float> {
  static __device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};

template <int MT, typename TIn, typename TOut>
__global__ void __launch_bounds__(NTHREADS)
conv_umma_kernel(const __grid_constant__ CUtensorMap wmap, const UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [W ring | A group buffers | barriers]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_ring = smem;
  uint8_t* a_bufs = w_ring + (size_t)p.stages * p.stage_bytes;
  const uint32_t a_buf_bytes = (uint32_t)p.PG * p.plane_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_bufs + (size_t)p.a_bufs * a_buf_bytes);
  uint64_t* full = bars;                    // [stages]
  uint64_t* empty = full + p.stages;        // [stages]
  uint64_t* a_full = empty + p.stages;      // [a_bufs]
  uint64_t* a_empty = a_full + p.a_bufs;    // [a_bufs]
  uint64_t* d_ready = a_empty + p.a_bufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_ready + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * (BM * MT);
  const int ntile = blockIdx.z;
  const uint32_t tapmask = p.tapmask ? p.tapmask[ntile] : 0xFFFFFFFFu;

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < p.a_bufs; ++s) {
      mbar_init(&a_full[s], 128);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(d_ready, 1);
    fence_barrier_init();
  }
  if (warp == 5) tcgen05_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  if (warp == 4 && lane == 0)
    asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int chunks_per_group = p.PG * 8 / p.KC;
  const int n_chunks = p.Cin / p.KC;

  if (warp == 4) {
    // ===== TMA producer: weights in (chunk, tap) order =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int chunk = 0; chunk < n_chunks; ++chunk) {
        for (int tap = 0; tap < p.K; ++tap) {
          if (!((tapmask >> tap) & 1u)) continue;
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], (uint32_t)p.stage_bytes);
          tma_load_2d(w_ring + (size_t)stage * p.stage_bytes, &wmap, &full[stage], chunk * p.KC,
                      tap * p.Cout + ntile * p.NT);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t w_layout = p.KC == 64 ? LAYOUT_SW128 : LAYOUT_SW64;
      const uint32_t w_sbo = 8u * (uint32_t)p.KC * 2u;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t started = 0;
      for (int g = 0; g < p.n_groups; ++g) {
        const int ab = g % p.a_bufs;
        mbar_wait(&a_full[ab], (uint32_t)(g / p.a_bufs) & 1u);
        tcgen05_fence_after();
        const uint32_t a_base = smem_u32(a_bufs + (size_t)ab * a_buf_bytes);
        const int c_end = min(n_chunks, (g + 1) * chunks_per_group);
        for (int chunk = g * chunks_per_group; chunk < c_end; ++chunk) {
          const int chunk_in_group = chunk - g * chunks_per_group;
          for (int tap = 0; tap < p.K; ++tap) {
            if (!((tapmask >> tap) & 1u)) continue;
            mbar_wait(&full[stage], phase);
            tcgen05_fence_after();
            const uint32_t w_base = smem_u32(w_ring + (size_t)stage * p.stage_bytes);
            const uint32_t row_shift = (uint32_t)(tap * p.dil) * 16u;
            for (int k16 = 0; k16 < p.KC / 16; ++k16) {
              const uint32_t plane = (uint32_t)(chunk_in_group * (p.KC / 8) + 2 * k16);
              const uint64_t bdesc = make_desc(w_base + k16 * 32, 0u, w_sbo, w_layout);
#pragma unroll
              for (int m = 0; m < MT; ++m) {
                const uint64_t adesc = make_desc(a_base + plane * p.plane_bytes + row_shift + m * (BM * 16),
                                                 (uint32_t)p.plane_bytes, 128u, LAYOUT_NONE);
                umma_f16(tmem_base + (uint32_t)(m * p.NT), adesc, bdesc, p.idesc, started);
              }
              started = 1;
            }
            tcgen05_commit(&empty[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        tcgen05_commit(&a_empty[ab]);
      }
      tcgen05_commit(d_ready);
    }
  } else {
    // ===== warps 0-3: stage the activation window group by group, then epilogue =====
    const int len = p.lens ? p.lens[b] : p.L;
    {
      const int rows = BM * MT + (p.K - 1) * p.dil;
      const TIn* xb = reinterpret_cast<const TIn*>(p.x) + (size_t)b * p.L * p.x_ld + p.x_coff;
      const float slope = p.in_slope;
      const int planes_total = p.Cin / 8;
      for (int g = 0; g < p.n_groups; ++g) {
        const int ab = g % p.a_bufs;
        mbar_wait(&a_empty[ab], ((uint32_t)(g / p.a_bufs) & 1u) ^ 1u);
        uint8_t* dst = a_bufs + (size_t)ab * a_buf_bytes;
        const int pl0 = g * p.PG;
        const int npl = min(p.PG, planes_total - pl0);
        const int nvec = rows * npl;
        for (int v = tid; v < nvec; v += 128) {
          const int r = v / npl, pl = v - r * npl;
          const int t = t0 - p.pad + r;
          uint4 q = make_uint4(0u, 0u, 0u, 0u);
          if (t >= 0 && t < p.L && (!p.in_mask || t < len)) {
            float f[8];
            VecIO<TIn>::load8(xb + (size_t)t * p.x_ld + (pl0 + pl) * 8, f);
            if (slope != 1.f) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = f[i] > 0.f ? f[i] : f[i] * slope;
            }
            __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
          }
          *reinterpret_cast<uint4*>(dst + (size_t)pl * p.plane_bytes + (size_t)r * 16) = q;
        }
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(&a_full[ab]);
      }
    }
    mbar_wait(d_ready, 0);
    tcgen05_fence_after();
    const int n0 = ntile * p.NT;
    const float* bbias = p.bbias ? p.bbias + (size_t)b * p.bbias_ld + n0 : nullptr;
    const float* bias = p.bias ? p.bias + n0 : nullptr;
#pragma unroll 1
    for (int m = 0; m < MT; ++m) {
      const int row = t0 + m * BM + warp * 32 + lane;
      const bool row_ok = row < p.L;
      const size_t grow = (size_t)b * p.L + (row_ok ? row : 0);
      TOut* yrow = reinterpret_cast<TOut*>(p.y) + grow * p.y_ld + p.y_coff + n0;
      const TOut* rrow = p.res ? reinterpret_cast<const TOut*>(p.res) + grow * p.res_ld + p.res_coff + n0 : nullptr;
      const bool zero_row = p.out_mask && row >= len;
#pragma unroll 1
      for (int c0 = 0; c0 < p.NT; c0 += 32) {
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(m * p.NT + c0), acc);
        if (!row_ok) continue;
#pragma unroll
        for (int q8 = 0; q8 < 4; ++q8) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[q8 * 8 + j]);
          const int c = c0 + q8 * 8;
          if (bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += __ldg(bias + c + j);
          }
          if (bbias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += __ldg(bbias + c + j);
          }
          if (rrow) {
            float r[8];
            VecIO<TOut>::load8(rrow + c, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += p.res_scale * r[j];
          }
          if (p.out_scale != 1.f) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] *= p.out_scale;
          }
          if (p.accumulate) {
            float o[8];
            VecIO<TOut>::load8(yrow + c, o);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += o[j];
          }
          if (p.act == ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.out_slope;
          } else if (p.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (zero_row) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
          }
          VecIO<TOut>::store8(yrow + c, v);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) tcgen05_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------ host side ----------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr; int cin, cout, k, kc, nt;
  bool operator<(const MapKey& o) const {
    return std::tie(ptr, cin, cout, k, kc, nt) < std::tie(o.ptr, o.cin, o.cout, o.k, o.kc, o.nt);
  }
};
std::map<MapKey, CUtensorMap> g_maps;
std::mutex g_maps_mu;

// weights [K][Cout][Cin] f16 viewed as a 2-D tensor {Cin, K*Cout}; box {KC, NT}
bool get_wmap(const void* w16, int cin, int cout, int k, int kc, int nt, CUtensorMap* out) {
  std::lock_guard<std::mutex> lk(g_maps_mu);
  MapKey key{w16, cin, cout, k, kc, nt};
  auto it = g_maps.find(key);
  if (it != g_maps.end()) {
    *out = it->second;
    return true;
  }
  auto enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)k * cout};
  cuuint64_t strides[1] = {(cuuint64_t)cin * 2};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)nt};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w16), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  g_maps[key] = m;
  *out = m;
  return true;
}

int pick_nt(int cout) {
  for (int nt : {256, 192, 128, 96, 64, 32})
    if (cout % nt == 0) return nt;
  return 0;
}

struct Plan {
  int MT, NT, KC, PG, n_groups, a_bufs, stages, stage_bytes, plane_bytes, tmem_cols;
  size_t smem;
};

// Choose the row-tile multiplicity, A grouping and weight-ring depth.  Preference: fit two
// CTAs per SM (<= 113 KB smem, <= 256 TMEM columns each) so one CTA's epilogue overlaps the
// other's MMAs; more row tiles per CTA amortise the weight stream from L2.
bool make_plan(const ConvArgs& a, Plan* out) {
  const int NT = pick_nt(a.Cout);
  if (!NT) return false;
  const int KC = a.Cin % 64 == 0 ? 64 : 32;
  const int planes = a.Cin / 8;
  const int PG = planes < GROUP_PLANES ? planes : GROUP_PLANES;
  if ((PG * 8) % KC) return false;
  const int n_groups = (planes + PG - 1) / PG;
  const int a_bufs = n_groups > 1 ? 2 : 1;
  const int stage_bytes = NT * KC * 2;
  const int n_tiles_rows = (a.L_out + BM - 1) / BM;
  const size_t fixed = 1024 + 512;
  static const int forced_mt = [] {
    const char* e = getenv("PG_UMMA_MT");   // tuning aid
    return e ? atoi(e) : 0;
  }();
  Plan best{};
  bool have = false;
  for (int pass = 0; pass < 2 && !have; ++pass) {
    const size_t budget = pass == 0 ? 113 * 1024 : 227 * 1024;
    const int tmem_budget = pass == 0 ? 256 : 512;
    for (int MT : {4, 2, 1}) {
      if (MT > 1 && (NT & (NT - 1))) continue;          // multi-tile only with power-of-two N
      if (MT * NT > tmem_budget) continue;
      if (forced_mt > 0 && MT != forced_mt && MT != 1) continue;
      const long ctas = (long)((n_tiles_rows + MT - 1) / MT) * a.B * (a.Cout / NT);
      if (MT > 1 && forced_mt <= 0 && ctas < 2 * 148) continue;   // do not starve the grid on short inputs
      const int rows = BM * MT + (a.K - 1) * a.dil;
      const int rows_alloc = (rows + 7) & ~7;
      const int plane_bytes = 16 * (rows_alloc + 1);
      const size_t a_bytes = (size_t)a_bufs * PG * plane_bytes;
      if (a_bytes + 2 * (size_t)stage_bytes + fixed > budget) continue;
      int stages = (int)((budget - fixed - a_bytes) / stage_bytes);
      if (stages > 6) stages = 6;
      int tm = 32;
      while (tm < MT * NT) tm <<= 1;
      best = Plan{MT, NT, KC, PG, n_groups, a_bufs, stages, stage_bytes, plane_bytes, tm,
                  fixed + a_bytes + (size_t)stages * stage_bytes};
      have = true;
      break;
    }
  }
  if (!have) return false;
  *out = best;
  return true;
}

template <int MT, typename TIn, typename TOut>
cudaError_t launch_t(const ConvArgs& a, const Plan& pl, cudaStream_t s) {
  CUtensorMap wmap;
  if (!get_wmap(a.w16, a.Cin, a.Cout, a.K, pl.KC, pl.NT, &wmap)) return cudaErrorNotSupported;
  UmmaParams p;
  p.x = a.x; p.x_ld = a.x_ld; p.x_coff = a.x_coff;
  p.res = a.res; p.res_ld = a.res_ld; p.res_coff = a.res_coff; p.res_scale = a.res_scale;
  p.y = a.y; p.y_ld = a.y_ld; p.y_coff = a.y_coff;
  p.bias = a.bias; p.bbias = a.bbias; p.bbias_ld = a.bbias_ld;
  p.lens = a.lens; p.in_mask = a.in_mask; p.out_mask = a.out_mask;
  p.tapmask = a.tapmask;
  p.B = a.B; p.L = a.L_out; p.Cin = a.Cin; p.Cout = a.Cout; p.K = a.K; p.dil = a.dil; p.pad = a.pad;
  p.NT = pl.NT;
  p.in_slope = a.in_slope; p.out_slope = a.out_slope; p.out_scale = a.out_scale;
  p.act = a.act; p.accumulate = a.accumulate;
  p.plane_bytes = pl.plane_bytes; p.KC = pl.KC; p.PG = pl.PG; p.n_groups = pl.n_groups;
  p.a_bufs = pl.a_bufs; p.stages = pl.stages; p.stage_bytes = pl.stage_bytes;
  p.tmem_cols = pl.tmem_cols;
  p.idesc = make_idesc(BM, pl.NT);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<MT, TIn, TOut>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  dim3 grid((a.L_out + BM * MT - 1) / (BM * MT), a.B, a.Cout / pl.NT);
  conv_umma_kernel<MT, TIn, TOut><<<grid, NTHREADS, pl.smem, s>>>(wmap, p);
  return cudaGetLastError();
}

template <typename TIn, typename TOut>
cudaError_t launch_mt(const ConvArgs& a, const Plan& pl, cudaStream_t s) {
  switch (pl.MT) {
    case 1: return launch_t<1, TIn, TOut>(a, pl, s);
    case 2: return launch_t<2, TIn, TOut>(a, pl, s);
    case 4: return launch_t<4, TIn, TOut>(a, pl, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace

int umma_pick_nt(int cout) { return pick_nt(cout); }

bool umma_conv_supported(const ConvArgs& a) {
  if (!a.w16) return false;
  if (a.Cin % 32 || a.Cin < 32) return false;
  if (a.x_ld % 8 || a.x_coff % 8 || a.y_ld % 8 || a.y_coff % 8) return false;
  if (a.res && (a.res_ld % 8 || a.res_coff % 8)) return false;
  if (a.L_in != a.L_out) return false;
  if (a.act == ACT_TANH) return false;
  if (a.K < 1 || a.K > 32) return false;
  if ((a.in_mask || a.out_mask) && !a.lens) return false;
  Plan pl;
  return make_plan(a, &pl);
}

cudaError_t launch_conv_umma(const ConvArgs& a, DType in_dt, DType out_dt, cudaStream_t s) {
  Plan pl;
  if (!umma_conv_supported(a) || !make_plan(a, &pl)) return cudaErrorInvalidValue;
  if (in_dt == DT_F16 && out_dt == DT_F16) return launch_mt<__half, __half>(a, pl, s);
  if (in_dt == DT_F32 && out_dt == DT_F32) return launch_mt<float, float>(a, pl, s);
  if (in_dt == DT_F32 && out_dt == DT_F16) return launch_mt<float, __half>(a, pl, s);
  return launch_mt<__half, float>(a, pl, s);
}

}  // namespace pg
