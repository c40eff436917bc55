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
#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>

#include "pg_common.cuh"
#include "pg_umma.cuh"

namespace pg {

namespace {

constexpr int BM = 128;              // rows per accumulator tile
constexpr int EPI_WARPS = 8, LOAD_WARPS = 8;
constexpr int LOAD_THREADS = LOAD_WARPS * 32, EPI_THREADS = EPI_WARPS * 32;
constexpr int TMA_WARP = EPI_WARPS + LOAD_WARPS, MMA_WARP = TMA_WARP + 1;
constexpr int NTHREADS = (MMA_WARP + 1) * 32;
constexpr int GROUP_PLANES = 16;     // 128 channels per activation-ring slot

using namespace umma;

struct UmmaParams {
  const void* x; int x_ld, x_coff;
  const void* res; int res_ld, res_coff; float res_scale;
  void* y; int y_ld, y_coff;
  const float* bias; const float* bbias; int bbias_ld;
  const int* lens; int in_mask, out_mask;
  const uint32_t* tapmask;     // [n_ntiles] bit t set = tap t has non-zero weights (nullable)
  int B, L;
  int Cin, Cout, K, dil, pad;
  int NT;                       // Cout tile (accumulator columns)
  int n_row_tiles, n_ntiles, total_tiles;
  float in_slope, out_slope, out_scale;
  int act, accumulate;
  int plane_bytes;              // A plane pitch (16 B * (rows_alloc + pad))
  int KC;                       // channels per weight stage: 64 (SW128) or 32 (SW64)
  int PG, n_groups;             // planes per A group, groups per tile
  int a_slots, a_slot_bytes;    // activation ring
  int stages, stage_bytes;      // weight ring (or, when resident, all K*chunks tiles of the layer)
  int resident;                 // 1: weights loaded once per CTA, no ring
  int split;                    // 1: two-term f16 operands (x = hi + lo), 3 MMAs per product: fp32-class accuracy
  int lo_off;                   // byte offset of the lo planes inside an activation slot
  int w_lo_rows;                // row offset (K*Cout) of the lo weight copy in the tensor map
  int rotate;                   // ring mode: per-row-tile rotation of the (chunk, tap) walk
  int tmem_cols;
  uint32_t idesc;
  int debug;   // PG_UMMA_DEBUG bitmask (timing experiments only): 1 no A loads, 2 no epilogue I/O, 4 no MMA
};

template <typename T> struct VecIO;
template <> struct VecIO<__half> {
  static constexpr int RAW = 1;   // uint4 per 8 channels
  static __device__ __forceinline__ void load_raw(const __half* p, uint4 (&q)[1]) {
    q[0] = *reinterpret_cast<const uint4*>(p);
  }
  static __device__ __forceinline__ void unpack(const uint4 (&q)[1], float (&v)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&q[0]);
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
  static constexpr int RAW = 2;
  static __device__ __forceinline__ void load_raw(const float* p, uint4 (&q)[2]) {
    q[0] = *reinterpret_cast<const uint4*>(p);
    q[1] = *reinterpret_cast<const uint4*>(p + 4);
  }
  static __device__ __forceinline__ void unpack(const uint4 (&q)[2], float (&v)[8]) {
    v[0] = __uint_as_float(q[0].x); v[1] = __uint_as_float(q[0].y);
    v[2] = __uint_as_float(q[0].z); v[3] = __uint_as_float(q[0].w);
    v[4] = __uint_as_float(q[1].x); v[5] = __uint_as_float(q[1].y);
    v[6] = __uint_as_float(q[1].z); v[7] = __uint_as_float(q[1].w);
  }
  static __device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};

// PG_UMMA_DEBUG & 8: CTA 0 records (role, event, tile, globaltimer) -- timing experiments only
__device__ unsigned long long g_trace[8192];
__device__ unsigned int g_trace_n;
__device__ __forceinline__ void trace(const UmmaParams& p, int role, int ev, int tile) {
  if ((p.debug & 8) && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned int i = atomicAdd(&g_trace_n, 1u);
    if (i < 8192) g_trace[i] = (t << 20) | ((unsigned long long)role << 16) | ((unsigned long long)ev << 12) | (unsigned)(tile & 0xFFF);
  }
}

struct TileCoord { int rt, b, ntile; };
__device__ __forceinline__ TileCoord decode_tile(int id, const UmmaParams& p) {
  TileCoord c;
  c.rt = id % p.n_row_tiles;
  const int tmp = id / p.n_row_tiles;
  c.b = tmp % p.B;
  c.ntile = tmp / p.B;
  return c;
}

// Persistent, warp-specialised: every CTA walks tiles id = blockIdx.x, +gridDim.x, ...
//   warps 0-7   epilogue  (TMEM lane quarter = warp % 4, column half = warp / 4)
//   warps 8-15  A loaders (global -> f16 planes in the activation ring)
//   warp  16    TMA producer for the weight ring
//   warp  17    TMEM owner + MMA issuer
// Rings: activation slots (a_full/a_empty), weight stages (full/empty), two TMEM
// accumulator sets (acc_full/acc_empty) so the epilogue of tile i overlaps the MMAs of i+1.
template <int MT, typename TIn, typename TOut>
__global__ void __launch_bounds__(NTHREADS)
conv_umma_kernel(const __grid_constant__ CUtensorMap wmap, const UmmaParams p) {
  pdl_launch_dependents();     // the next kernel of the stream may launch and run its prologue (pg_common.cuh)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_ring = smem;
  uint8_t* a_ring = w_ring + (size_t)p.stages * p.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_ring + (size_t)p.a_slots * p.a_slot_bytes);
  uint64_t* full = bars;                      // [stages]
  uint64_t* empty = full + p.stages;          // [stages]
  uint64_t* a_full = empty + p.stages;        // [a_slots]
  uint64_t* a_empty = a_full + p.a_slots;     // [a_slots]
  uint64_t* acc_full = a_empty + p.a_slots;   // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint64_t* w_ready = acc_empty + 2;          // resident weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_ready + 1);
  // per-epilogue-warp transpose buffers (f16 output): 32 rows x 64 B, 80 B pitch (conflict-free)
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(tmem_slot + 4) + 127) & ~uintptr_t(127));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < p.a_slots; ++s) {
      mbar_init(&a_full[s], LOAD_THREADS);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], EPI_THREADS);
    }
    mbar_init(w_ready, 1);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tcgen05_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  if (warp == TMA_WARP && lane == 0)
    asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) trace(p, 4, 0, 0);
  // activations in global memory belong to the predecessor until it has completed; the weight producer and the MMA
  // issuer touch constants / shared memory / TMEM only and run ahead
  if (warp != TMA_WARP && warp != MMA_WARP) pdl_wait();

  const int chunks_per_group = p.PG * 8 / p.KC;
  const int n_chunks = p.Cin / p.KC;
  const int planes_total = p.Cin / 8;

  if (warp == TMA_WARP) {
    // ===== TMA producer: weights in (tile, chunk, tap) order =====
    if (p.resident) {
      // the whole layer's weights fit: one bulk load per (chunk, tap) tile, once per CTA
      if (elect_one()) {
        trace(p, 0, 0, 0);
        mbar_expect_tx(w_ready, (uint32_t)(n_chunks * p.K * p.stage_bytes));
        const int tile_b = p.split ? p.stage_bytes / 2 : p.stage_bytes;
        for (int chunk = 0; chunk < n_chunks; ++chunk)
          for (int tap = 0; tap < p.K; ++tap) {
            uint8_t* dstw = w_ring + (size_t)(chunk * p.K + tap) * p.stage_bytes;
            tma_load_2d(dstw, &wmap, w_ready, chunk * p.KC, tap * p.Cout);
            if (p.split) tma_load_2d(dstw + tile_b, &wmap, w_ready, chunk * p.KC, p.w_lo_rows + tap * p.Cout);
          }
      }
      __syncwarp();
    } else {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(tile, p);
        const uint32_t tapmask = p.tapmask ? p.tapmask[tc.ntile] : 0xFFFFFFFFu;
        // Every CTA streams the same weight tiles: start each tile's (chunk, tap) walk at a
        // row-tile dependent offset so the SMs do not hammer the same L2 lines in lock-step.
        for (int g = 0; g < p.n_groups; ++g) {
          const int c_begin = g * chunks_per_group;
          const int n_it = (min(n_chunks, c_begin + chunks_per_group) - c_begin) * p.K;
          const int rot = p.rotate ? (tc.rt * 5) % n_it : 0;
          for (int i = 0; i < n_it; ++i) {
            int idx = i + rot;
            if (idx >= n_it) idx -= n_it;
            const int chunk = c_begin + idx / p.K, tap = idx % p.K;
            if (!((tapmask >> tap) & 1u)) continue;
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&full[stage], (uint32_t)p.stage_bytes);
              tma_load_2d(w_ring + (size_t)stage * p.stage_bytes, &wmap, &full[stage], chunk * p.KC,
                          tap * p.Cout + tc.ntile * p.NT);
              if (p.split)
                tma_load_2d(w_ring + (size_t)stage * p.stage_bytes + p.stage_bytes / 2, &wmap, &full[stage],
                            chunk * p.KC, p.w_lo_rows + tap * p.Cout + tc.ntile * p.NT);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===== MMA issuer: the whole warp walks the loops, one elected lane issues =====
    const uint32_t w_layout = p.KC == 64 ? LAYOUT_SW128 : LAYOUT_SW64;
    const uint32_t w_sbo = 8u * (uint32_t)p.KC * 2u;
    // descriptor templates: start address (low 14 bits, 16-byte units) is OR-ed in per MMA
    const uint64_t adesc0 = make_desc(0u, (uint32_t)p.plane_bytes, 128u, LAYOUT_NONE);
    const uint64_t bdesc0 = make_desc(0u, 0u, w_sbo, w_layout);
    const uint32_t plane_units = (uint32_t)p.plane_bytes >> 4;
    const int kc16 = p.KC / 16, kc8 = p.KC / 8;
    const uint32_t nt = (uint32_t)p.NT, idesc = p.idesc;
    const bool do_mma = !(p.debug & 4);
    int stage = 0;
    uint32_t phase = 0, a_cnt = 0, t_cnt = 0;
    if (p.resident) {
      mbar_wait(w_ready, 0);
      tcgen05_fence_after();
      trace(p, 1, 0, 0);
    }
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t_cnt) {
      const TileCoord tc = decode_tile(tile, p);
      const uint32_t tapmask = p.tapmask ? p.tapmask[tc.ntile] : 0xFFFFFFFFu;
      const uint32_t accb = t_cnt & 1u;
      mbar_wait(&acc_empty[accb], ((t_cnt >> 1) & 1u) ^ 1u);
      tcgen05_fence_after();
      if (lane == 0) trace(p, 1, 1, (int)t_cnt);
      const uint32_t d_base = tmem_base + accb * (uint32_t)MT * nt;
      uint32_t started = 0;
      for (int g = 0; g < p.n_groups; ++g, ++a_cnt) {
        const uint32_t slot = a_cnt % (uint32_t)p.a_slots;
        mbar_wait(&a_full[slot], (a_cnt / (uint32_t)p.a_slots) & 1u);
        tcgen05_fence_after();
        if (lane == 0) trace(p, 1, 2, (int)t_cnt);
        const uint32_t a_units = smem_u32(a_ring + (size_t)slot * p.a_slot_bytes) >> 4;
        const int c_begin = g * chunks_per_group;
        const int n_it = (min(n_chunks, c_begin + chunks_per_group) - c_begin) * p.K;
        const int rot = p.rotate ? (tc.rt * 5) % n_it : 0;
        for (int i = 0; i < n_it; ++i) {
          int idx = i + rot;
          if (idx >= n_it) idx -= n_it;
          const int chunk_in_group = idx / p.K, tap = idx - chunk_in_group * p.K;
          const int chunk = c_begin + chunk_in_group;
          if (!((tapmask >> tap) & 1u)) continue;
          if (!p.resident) {
            mbar_wait(&full[stage], phase);
            tcgen05_fence_after();
          }
          const uint32_t w_units =
              smem_u32(w_ring + (size_t)(p.resident ? chunk * p.K + tap : stage) * p.stage_bytes) >> 4;
          const uint32_t a_tap = a_units + (uint32_t)(chunk_in_group * kc8) * plane_units + (uint32_t)(tap * p.dil);
          if (elect_one()) {
            if (do_mma) {
              for (int k16 = 0; k16 < kc16; ++k16) {
                const uint64_t bdesc = bdesc0 | (uint64_t)(w_units + 2u * k16);
                const uint32_t a_k = a_tap + 2u * (uint32_t)k16 * plane_units;
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                  const uint64_t adesc = adesc0 | (uint64_t)(a_k + (uint32_t)m * (BM * 16 / 16));
                  umma_f16(d_base + (uint32_t)m * nt, adesc, bdesc, idesc, started | (uint32_t)k16);
                  if (p.split) {   // + lo(A)*hi(B) + hi(A)*lo(B)
                    const uint64_t adesc_lo = adesc0 | (uint64_t)(a_k + (uint32_t)m * (BM * 16 / 16) + ((uint32_t)p.lo_off >> 4));
                    const uint64_t bdesc_lo = bdesc0 | (uint64_t)(w_units + 2u * k16 + ((uint32_t)p.stage_bytes >> 5));
                    umma_f16(d_base + (uint32_t)m * nt, adesc_lo, bdesc, idesc, 1u);
                    umma_f16(d_base + (uint32_t)m * nt, adesc, bdesc_lo, idesc, 1u);
                  }
                }
              }
            }
            if (!p.resident) tcgen05_commit(&empty[stage]);
          }
          __syncwarp();
          started = 1;
          if (!p.resident) {
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        if (elect_one()) tcgen05_commit(&a_empty[slot]);
        __syncwarp();
      }
      if (elect_one()) tcgen05_commit(&acc_full[accb]);
      __syncwarp();
      if (lane == 0) trace(p, 1, 3, (int)t_cnt);
    }
  } else if (warp >= EPI_WARPS) {
    // ===== A loaders: stage (tile, group) windows into the activation ring =====
    const int ltid = tid - EPI_THREADS;
    const int rows = BM * MT + (p.K - 1) * p.dil;
    const float slope = p.in_slope;
    // vectors (8 channels each) in flight per thread.  (Round 3 tried U = 8 for fp32 inputs so that a 128-row x
    // 16-plane group is one round trip instead of two: the 64 staging registers spill under the 96-register cap of
    // this 18-warp CTA and the step got 2 % slower -- tools/umma_trace.py shows the role timeline.)
    constexpr int U = VecIO<TIn>::RAW == 1 ? 8 : 4;
    uint32_t a_cnt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(tile, p);
      const int t0 = tc.rt * (BM * MT);
      const int len = p.lens ? p.lens[tc.b] : p.L;
      const int t_hi = p.in_mask ? min(p.L, len) : p.L;
      const TIn* xb = reinterpret_cast<const TIn*>(p.x) + (size_t)tc.b * p.L * p.x_ld + p.x_coff;
      for (int g = 0; g < p.n_groups; ++g, ++a_cnt) {
        const uint32_t slot = a_cnt % (uint32_t)p.a_slots;
        mbar_wait(&a_empty[slot], ((a_cnt / (uint32_t)p.a_slots) & 1u) ^ 1u);
        if (ltid == 0) trace(p, 2, 1, (int)a_cnt);
        uint8_t* dst = a_ring + (size_t)slot * p.a_slot_bytes;
        const int pl0 = g * p.PG;
        const int npl = min(p.PG, planes_total - pl0);
        const int nvec = rows * npl;
        const int npl_shift = (npl & (npl - 1)) == 0 ? 31 - __clz(npl) : -1;
        for (int v0 = ltid; v0 < nvec; v0 += LOAD_THREADS * U) {
          uint4 raw[U][VecIO<TIn>::RAW];
          int rr[U], pp[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int v = v0 + u * LOAD_THREADS;
            rr[u] = -1;
            if (v < nvec) {
              const int r = npl_shift >= 0 ? (v >> npl_shift) : v / npl;
              const int pl = v - r * npl;
              rr[u] = r;
              pp[u] = pl;
              const int t = t0 - p.pad + r;
              if (t >= 0 && t < t_hi && !(p.debug & 1)) {
                VecIO<TIn>::load_raw(xb + (size_t)t * p.x_ld + (pl0 + pl) * 8, raw[u]);
              } else {
#pragma unroll
                for (int i = 0; i < VecIO<TIn>::RAW; ++i) raw[u][i] = make_uint4(0u, 0u, 0u, 0u);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (rr[u] < 0) continue;
            float f[8];
            VecIO<TIn>::unpack(raw[u], f);
            if (slope != 1.f) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = f[i] > 0.f ? f[i] : f[i] * slope;
            }
            uint4 q;
            __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
            *reinterpret_cast<uint4*>(dst + (size_t)pp[u] * p.plane_bytes + (size_t)rr[u] * 16) = q;
            if (p.split) {
              uint4 ql;
              __half2* hl = reinterpret_cast<__half2*>(&ql);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 back = __half22float2(h[i]);
                hl[i] = __floats2half2_rn(f[2 * i] - back.x, f[2 * i + 1] - back.y);
              }
              *reinterpret_cast<uint4*>(dst + p.lo_off + (size_t)pp[u] * p.plane_bytes + (size_t)rr[u] * 16) = ql;
            }
          }
        }
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(&a_full[slot]);
        if (ltid == 0) trace(p, 2, 2, (int)a_cnt);
      }
    }
  } else {
    // ===== epilogue warps: quarter q = warp % 4 owns TMEM lanes 32q..32q+31 (one output row per
    // lane); the two warps of a quarter split the NT columns in 32-column blocks (even / odd).
    // f16 outputs go through a per-warp shared-memory transpose so that every global access is
    // coalesced (8 rows x 64 B per instruction instead of 32 rows x 16 B).
    constexpr int RV = VecIO<TOut>::RAW;   // uint4 per 8 output channels
    constexpr int BV = 4 * RV;             // uint4 per 32-column block of one row
    constexpr bool STAGED = RV == 1;
    constexpr int EP = STAGED ? 80 : 144;  // staging row pitch in bytes (f16: 64 B rows, f32: 128 B rows; bank-staggered)
    const int quarter = warp & 3, half = warp >> 2;
    const int n_blocks = p.NT / 32;        // 32-column blocks in the tile
    uint8_t* stg = epi_stage + (size_t)warp * (32 * EP);
    const int crow = lane >> 2, cchunk = lane & 3;   // coalesced mapping: rows crow + 8i, 16-byte chunk
    uint32_t t_cnt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t_cnt) {
      const TileCoord tc = decode_tile(tile, p);
      const int t0 = tc.rt * (BM * MT);
      const int len = p.lens ? p.lens[tc.b] : p.L;
      const int n0 = tc.ntile * p.NT;
      const float* bbias = p.bbias ? p.bbias + (size_t)tc.b * p.bbias_ld + n0 : nullptr;
      const float* bias = p.bias ? p.bias + n0 : nullptr;
      const uint32_t accb = t_cnt & 1u;
      bool waited = false;
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int wrow0 = t0 + m * BM + quarter * 32;     // first row of this warp
        const int row = wrow0 + lane;
        const bool row_ok = row < p.L;
        const size_t grow = (size_t)tc.b * p.L + (row_ok ? row : 0);
        TOut* yrow = reinterpret_cast<TOut*>(p.y) + grow * p.y_ld + p.y_coff + n0;
        const TOut* rrow = p.res ? reinterpret_cast<const TOut*>(p.res) + grow * p.res_ld + p.res_coff + n0 : nullptr;
        const bool zero_row = p.out_mask && row >= len;
        // coalesced-mapping bases (row wrow0 + crow + 8i)
        const TOut* rco = p.res ? reinterpret_cast<const TOut*>(p.res) + ((size_t)tc.b * p.L + wrow0) * p.res_ld +
                                      p.res_coff + n0 + cchunk * 8 : nullptr;
        TOut* yco = reinterpret_cast<TOut*>(p.y) + ((size_t)tc.b * p.L + wrow0) * p.y_ld + p.y_coff + n0 + cchunk * 8;
        uint4 r0[BV];
        uint4 r1[RV == 1 ? BV : 1];
        auto fetch = [&](int blk, uint4 (&dst)[BV]) {
          if (!p.res || blk >= n_blocks || (p.debug & 2)) return;
          if constexpr (STAGED) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = crow + 8 * i;
              dst[i] = make_uint4(0u, 0u, 0u, 0u);
              if (wrow0 + r < p.L)
                dst[i] = *reinterpret_cast<const uint4*>(rco + (size_t)r * p.res_ld + blk * 32);
            }
          } else {
            if (row_ok) {
#pragma unroll
              for (int g8 = 0; g8 < 4; ++g8) {
                uint4 tmp[RV];
                VecIO<TOut>::load_raw(rrow + blk * 32 + g8 * 8, tmp);
#pragma unroll
                for (int i = 0; i < RV; ++i) dst[g8 * RV + i] = tmp[i];
              }
            }
          }
        };
        fetch(half, r0);
        if constexpr (RV == 1) fetch(half + 2, r1);
        if (!waited) {
          mbar_wait(&acc_full[accb], (t_cnt >> 1) & 1u);
          tcgen05_fence_after();
          waited = true;
          if (tid == 0) trace(p, 3, 1, (int)t_cnt);
        }
        auto process = [&](int blk, uint4 (&rb)[BV]) {
          const int c0 = blk * 32;
          if constexpr (STAGED) {
            if (p.res) {   // coalesced residual -> staging -> per-row registers
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i)
                *reinterpret_cast<uint4*>(stg + (crow + 8 * i) * EP + cchunk * 16) = rb[i];
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) rb[i] = *reinterpret_cast<const uint4*>(stg + lane * EP + i * 16);
            }
          }
          uint32_t acc[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + accb * (uint32_t)(MT * p.NT) +
                        (uint32_t)(m * p.NT + c0), acc);
          if constexpr (STAGED) __syncwarp();   // staging reads done before it is rewritten
#pragma unroll
          for (int q8 = 0; q8 < 4; ++q8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[q8 * 8 + j]);
            const int c = c0 + q8 * 8;
            if (bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (bbias) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] += __ldg(bbias + c + j);
            }
            if (p.res) {
              uint4 tmp[RV];
#pragma unroll
              for (int i = 0; i < RV; ++i) tmp[i] = rb[q8 * RV + i];
              float r[8];
              VecIO<TOut>::unpack(tmp, r);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] += p.res_scale * r[j];
            }
            if (p.out_scale != 1.f) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] *= p.out_scale;
            }
            if (p.accumulate && row_ok) {
              uint4 oraw[RV];
              float o[8];
              VecIO<TOut>::load_raw(yrow + c, oraw);
              VecIO<TOut>::unpack(oraw, o);
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
            if constexpr (STAGED) {
              uint4 q;
              __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
              for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
              *reinterpret_cast<uint4*>(stg + lane * EP + q8 * 16) = q;
            } else if (p.act == ACT_GATE) {
              // WaveNet gate on (tanh, sigmoid) column pairs: 8 accumulator columns -> 4 outputs (same arithmetic as
              // the stand-alone gate kernel)
              float g[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) g[i] = tanhf(v[2 * i]) * (1.f / (1.f + __expf(-v[2 * i + 1])));
              *reinterpret_cast<float4*>(stg + lane * EP + q8 * 16) = make_float4(g[0], g[1], g[2], g[3]);
            } else {
              // f32 outputs: stage the row's 32 B so the block leaves as full 128 B rows (below)
              *reinterpret_cast<float4*>(stg + lane * EP + q8 * 32) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(stg + lane * EP + q8 * 32 + 16) = make_float4(v[4], v[5], v[6], v[7]);
            }
          }
          if constexpr (!STAGED) {
            __syncwarp();
            if (p.act == ACT_GATE) {      // 16 outputs (64 B) per row and block
              if (!(p.debug & 2)) {
                TOut* yb = reinterpret_cast<TOut*>(p.y) + ((size_t)tc.b * p.L + wrow0) * p.y_ld + p.y_coff + (n0 + c0) / 2;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int id = i * 32 + lane, r = id >> 2, c16 = id & 3;
                  if (wrow0 + r < p.L)
                    *reinterpret_cast<float4*>(yb + (size_t)r * p.y_ld + c16 * 4) =
                        *reinterpret_cast<const float4*>(stg + r * EP + c16 * 16);
                }
              }
            } else
            if (!(p.debug & 2)) {
              TOut* yb = reinterpret_cast<TOut*>(p.y) + ((size_t)tc.b * p.L + wrow0) * p.y_ld + p.y_coff + n0 + c0;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int id = i * 32 + lane, r = id >> 3, c16 = id & 7;
                if (wrow0 + r < p.L)
                  *reinterpret_cast<float4*>(yb + (size_t)r * p.y_ld + c16 * 4) =
                      *reinterpret_cast<const float4*>(stg + r * EP + c16 * 16);
              }
            }
            __syncwarp();   // staging is rewritten by the next block
          }
          if constexpr (STAGED) {
            __syncwarp();
            if (!(p.debug & 2)) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int r = crow + 8 * i;
                if (wrow0 + r < p.L)
                  *reinterpret_cast<uint4*>(yco + (size_t)r * p.y_ld + c0) =
                      *reinterpret_cast<const uint4*>(stg + r * EP + cchunk * 16);
              }
            }
          }
        };
        if constexpr (RV == 1) {
#pragma unroll 1
          for (int blk = half; blk < n_blocks; blk += 4) {
            process(blk, r0);
            fetch(blk + 4, r0);
            if (blk + 2 < n_blocks) {
              process(blk + 2, r1);
              fetch(blk + 6, r1);
            }
          }
        } else {   // f32 residual rows are twice as wide: one block in flight
#pragma unroll 1
          for (int blk = half; blk < n_blocks; blk += 2) {
            process(blk, r0);
            fetch(blk + 2, r0);
          }
        }
      }
      if (!waited) {   // this warp had no columns (NT == 32, half == 1): still track the phase
        mbar_wait(&acc_full[accb], (t_cnt >> 1) & 1u);
        tcgen05_fence_after();
      }
      tcgen05_fence_before();
      mbar_arrive(&acc_empty[accb]);
      if (tid == 0) trace(p, 3, 2, (int)t_cnt);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tcgen05_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  if (tid == 0) trace(p, 4, 1, 0);
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
  int MT, NT, KC, PG, n_groups, a_slots, a_slot_bytes, stages, stage_bytes, plane_bytes, tmem_cols;
  int ctas_per_sm, resident;
  size_t smem;
  int lo_off;
};

int env_int(const char* name) {
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
}

int num_sms() {
  static std::atomic<int> cache[PG_MAX_DEVICES] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const bool tracked = dev >= 0 && dev < PG_MAX_DEVICES;
  int v = tracked ? cache[dev].load(std::memory_order_relaxed) : 0;
  if (v > 0) return v;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 1;
  if (tracked) cache[dev].store(v, std::memory_order_relaxed);
  return v;
}

// Choose row-tile multiplicity, ring depths and CTAs per SM.  Narrow layers (little MMA work per
// tile) want several resident CTAs to hide the staging / epilogue latency; wide layers want deep
// rings in one CTA.  TMEM holds two accumulator sets per CTA (epilogue / MMA overlap).
bool make_plan_nt(const ConvArgs& a, int NT, Plan* out);

// Split precision doubles every operand tile: walk down the Cout tile sizes until the rings fit.
bool make_plan(const ConvArgs& a, Plan* out) {
  if (!a.split) {
    const int NT = pick_nt(a.Cout);
    return NT && make_plan_nt(a, NT, out);
  }
  for (int nt : {128, 96, 64, 32})
    if (a.Cout % nt == 0 && make_plan_nt(a, nt, out)) return true;
  return false;
}

bool make_plan_nt(const ConvArgs& a, int NT, Plan* out) {
  const int KC = a.Cin % 64 == 0 ? 64 : 32;
  const int planes = a.Cin / 8;
  int PG = planes < GROUP_PLANES ? planes : GROUP_PLANES;
  if ((PG * 8) % KC) PG = planes;   // e.g. Cin = 96 with 32-channel chunks: one group
  if ((PG * 8) % KC) return false;
  const int n_groups = (planes + PG - 1) / PG;
  const int stage_bytes = NT * KC * 2 * (a.split ? 2 : 1);
  static const int forced_mt = env_int("PG_UMMA_MT"), no_resident = env_int("PG_UMMA_NORESIDENT");
  const size_t fixed = 1024 + 1024 + EPI_WARPS * 32 * (a.out_f32 ? 144 : 80) + 256;   // align slack, barriers, epilogue staging
  const size_t budget = (size_t)226 * 1024;
  const int n_chunks = a.Cin / KC;
  const size_t w_all = (size_t)n_chunks * a.K * stage_bytes;
  const bool can_reside = !no_resident && a.Cout == NT && !a.tapmask;
  const int n_rows128 = (a.L_out + BM - 1) / BM;
  auto geometry = [&](int MT, int* plane_bytes, int* a_slot_bytes) {
    const int rows = BM * MT + (a.K - 1) * a.dil;
    const int rows_alloc = (rows + 7) & ~7;
    const int row_pad = PG >= 8 ? 1 : (PG == 4 ? 2 : 1);   // keep plane pitches on distinct banks
    *plane_bytes = 16 * (rows_alloc + row_pad);
    *a_slot_bytes = ((PG * *plane_bytes + 127) & ~127) * (a.split ? 2 : 1);
  };
  // Candidate row-tile multiplicities, largest first: more rows per tile amortise the barrier
  // round trips and (ring mode) the weight stream from L2.  Two accumulator sets must fit TMEM.
  for (int pass = 0; pass < 2; ++pass) {
    const bool resident = pass == 0;
    if (resident && !can_reside) continue;
    for (int MT : {4, 2, 1}) {
      if (forced_mt > 0 && MT != forced_mt && MT != 1) continue;
      if (MT > 1 && (NT & (NT - 1))) continue;
      if (2 * MT * NT > 512) continue;
      if (MT > 1 && (long)((n_rows128 + MT - 1) / MT) * a.B * (a.Cout / NT) < 2L * num_sms()) continue;
      int plane_bytes, a_slot_bytes;
      geometry(MT, &plane_bytes, &a_slot_bytes);
      int tm = 32;
      while (tm < 2 * MT * NT) tm <<= 1;
      const int min_stages = a.split ? 2 : 3;
      const size_t w_bytes = resident ? w_all : min_stages * (size_t)stage_bytes;
      if (fixed + 2 * (size_t)a_slot_bytes + w_bytes > budget) continue;
      size_t left = budget - fixed - 2 * (size_t)a_slot_bytes - w_bytes;
      int a_slots = 2, stages = resident ? n_chunks * a.K : min_stages;
      if (!resident && env_int("PG_UMMA_STAGES") == 2) { stages = 2; left += stage_bytes; }
      static const int forced_stages = env_int("PG_UMMA_STAGES");
      const int max_stages = forced_stages > 0 ? forced_stages : 8;
      if (!resident) {   // ring: up to 8 stages, then deeper activation ring
        while (stages < max_stages && left >= (size_t)stage_bytes && stages < n_chunks * a.K) {
          ++stages;
          left -= stage_bytes;
        }
      }
      while (a_slots < 4 && left >= (size_t)a_slot_bytes) {
        ++a_slots;
        left -= a_slot_bytes;
      }
      *out = Plan{MT, NT, KC, PG, n_groups, a_slots, a_slot_bytes, stages, stage_bytes, plane_bytes, tm, 1,
                  resident ? 1 : 0,
                  fixed + (size_t)a_slots * a_slot_bytes + (size_t)stages * stage_bytes,
                  a.split ? a_slot_bytes / 2 : 0};
      return true;
    }
  }
  return false;
}

template <int MT, typename TIn, typename TOut>
cudaError_t launch_t(const ConvArgs& a, const Plan& pl, cudaStream_t s) {
  CUtensorMap wmap;
  if (!get_wmap(a.split ? a.w16s : a.w16, a.Cin, a.Cout, a.split ? 2 * a.K : a.K, pl.KC, pl.NT, &wmap))
    return cudaErrorNotSupported;
  UmmaParams p;
  p.split = a.split; p.lo_off = pl.lo_off; p.w_lo_rows = a.K * a.Cout;
  p.x = a.x; p.x_ld = a.x_ld; p.x_coff = a.x_coff;
  p.res = a.res; p.res_ld = a.res_ld; p.res_coff = a.res_coff; p.res_scale = a.res_scale;
  p.y = a.y; p.y_ld = a.y_ld; p.y_coff = a.y_coff;
  p.bias = a.bias; p.bbias = a.bbias; p.bbias_ld = a.bbias_ld;
  p.lens = a.lens; p.in_mask = a.in_mask; p.out_mask = a.out_mask;
  p.tapmask = a.tapmask;
  p.B = a.B; p.L = a.L_out; p.Cin = a.Cin; p.Cout = a.Cout; p.K = a.K; p.dil = a.dil; p.pad = a.pad;
  p.NT = pl.NT;
  p.n_row_tiles = (a.L_out + BM * MT - 1) / (BM * MT);
  p.n_ntiles = a.Cout / pl.NT;
  p.total_tiles = p.n_row_tiles * a.B * p.n_ntiles;
  p.in_slope = a.in_slope; p.out_slope = a.out_slope; p.out_scale = a.out_scale;
  p.act = a.act; p.accumulate = a.accumulate;
  p.plane_bytes = pl.plane_bytes; p.KC = pl.KC; p.PG = pl.PG; p.n_groups = pl.n_groups;
  p.a_slots = pl.a_slots; p.a_slot_bytes = pl.a_slot_bytes;
  p.stages = pl.stages; p.stage_bytes = pl.stage_bytes; p.resident = pl.resident;
  p.tmem_cols = pl.tmem_cols;
  p.idesc = make_idesc(BM, pl.NT);
  static const int dbg = env_int("PG_UMMA_DEBUG"), norot = env_int("PG_UMMA_NOROTATE");
  p.debug = dbg;
  p.rotate = (!pl.resident && !norot) ? 1 : 0;
  static DeviceOnce once;
  if (cudaError_t e = ensure_dyn_smem(conv_umma_kernel<MT, TIn, TOut>, once, 227 * 1024)) return e;
  int grid = num_sms() * pl.ctas_per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  if (p.debug & 8) {
    unsigned int zero = 0;
    cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(zero));
  }
  if (cudaError_t e = launch_pdl(conv_umma_kernel<MT, TIn, TOut>, dim3(grid), dim3(NTHREADS), pl.smem, s, wmap, p)) return e;
  if (p.debug & 8) {
    cudaDeviceSynchronize();
    static unsigned long long host[8192];
    unsigned int n = 0;
    cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(n));
    cudaMemcpyFromSymbol(host, g_trace, sizeof(host));
    if (n > 8192) n = 8192;
    unsigned long long t0 = ~0ull;
    for (unsigned i = 0; i < n; ++i) t0 = host[i] >> 20 < t0 ? host[i] >> 20 : t0;
    fprintf(stderr, "TRACE MT=%d NT=%d resident=%d a_slots=%d stages=%d grid=%d tiles=%d smem=%zu\n", MT, pl.NT,
            pl.resident, pl.a_slots, pl.stages, grid, p.total_tiles, pl.smem);
    for (unsigned i = 0; i < n; ++i)
      fprintf(stderr, "TR %llu role=%llu ev=%llu tile=%llu\n", (host[i] >> 20) - t0, (host[i] >> 16) & 15,
              (host[i] >> 12) & 15, host[i] & 0xFFF);
  }
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

bool get_weight_map(const void* w16, int cin, int cout, int k, int kc, int nt, CUtensorMap* out) {
  return get_wmap(w16, cin, cout, k, kc, nt, out);
}
int device_sm_count() { return num_sms(); }
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("PG_PDL"); return !e || atoi(e) != 0; }();
  return on;
}

int umma_pick_nt(int cout) { return pick_nt(cout); }

bool umma_conv_supported(const ConvArgs& a) {
  if (!a.w16) return false;
  if (a.split && (!a.w16s || a.tapmask)) return false;
  if (a.Cin % 32 || a.Cin < 32) return false;
  if (a.x_ld % 8 || a.x_coff % 8 || a.y_ld % 8 || a.y_coff % 8) return false;
  if (a.res && (a.res_ld % 8 || a.res_coff % 8)) return false;
  if (a.L_in != a.L_out) return false;
  if (a.act == ACT_TANH) return false;
  if (a.act == ACT_GATE && (!a.out_f32 || a.accumulate || a.res || a.out_mask || a.Cout % 64)) return false;
  if (a.K < 1 || a.K > 32) return false;
  if ((a.in_mask || a.out_mask) && !a.lens) return false;
  Plan pl;
  return make_plan(a, &pl);
}

cudaError_t launch_conv_umma(const ConvArgs& a_in, DType in_dt, DType out_dt, cudaStream_t s) {
  ConvArgs a = a_in;
  a.out_f32 = out_dt == DT_F32 ? 1 : 0;
  Plan pl;
  if (!umma_conv_supported(a) || !make_plan(a, &pl)) return cudaErrorInvalidValue;
  if (in_dt == DT_F16 && out_dt == DT_F16) return launch_mt<__half, __half>(a, pl, s);
  if (in_dt == DT_F32 && out_dt == DT_F32) return launch_mt<float, float>(a, pl, s);
  if (in_dt == DT_F32 && out_dt == DT_F16) return launch_mt<float, __half>(a, pl, s);
  return launch_mt<__half, float>(a, pl, s);
}

}  // namespace pg
