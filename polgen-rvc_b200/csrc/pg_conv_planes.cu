// tcgen05 / TMEM implicit-GEMM Conv1d over CHANNEL-PLANE activations (sm_100a) -- the decoder's
// conv kernel: ResBlock convs (reference rvc/lib/algorithm/residuals.py:45-53), the polyphase
// form of the ConvTranspose1d upsamplers (nsf.py:80-91) and conv_pre (nsf.py:64-66).
//
// Activation layout in HBM ("planes"): f16 [B][C/8][L][8] -- for every group of 8 channels one
// contiguous [L][8] array, 16 bytes per time row.  This IS the UMMA no-swizzle K-major canonical
// layout (8-row x 16-byte core matrices, SBO = 128 B, LBO = plane pitch), so
//   * the A operand window of a tile (MT*128 + (K-1)*dil rows) is fetched by one 1-D bulk copy
//     (cp.async.bulk -> UBLKCP) per plane straight into shared memory: no register staging, no
//     conversion, deep bytes-in-flight, completion on an mbarrier;
//   * a tap / dilation shift is `start_address += shift*16` in the descriptor (zero-copy implicit
//     im2col), every tap and every one of the MT row tiles re-reads the same window;
//   * the epilogue's per-lane 16-byte accesses (one TMEM lane = one time row) are contiguous across
//     the warp: residual reads and output writes are fully coalesced without a transpose.
// Conv inputs are stored post-activation ("L-form": lrelu(x, slope) in f16); the raw value a
// residual needs is recovered in registers (x = v < 0 ? v/slope : v), which is exact to the same
// 2^-11 as storing x itself.  Sequence ends / masks are handled by zero-filling the rows of the
// window that fall outside [0, t_hi) in the loader warp; there are no guard rows in HBM.
//
// D[M = 128 time rows, N = Cout tile] += A * B per tap and per 16-channel K slice, f16 operands,
// fp32 accumulation in TMEM.  B = weights [tap][Cout][Cin] f16, streamed by TMA (128B swizzle)
// through an mbarrier ring, or loaded once per CTA when the layer fits in shared memory.
// Persistent CTAs (one per SM), warp-specialised:
//   warp 0      A loader (bulk copies + boundary zero fill)
//   warp 1      weight TMA producer
//   warp 2      TMEM owner + MMA issuer
//   warps 4-11  epilogue: TMEM -> registers -> bias / residual / scale / accumulate / leaky-ReLU
//               -> global.  Two accumulator sets so the epilogue of tile i overlaps the MMAs of i+1.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>

#include "pg_common.cuh"
#include "pg_umma.cuh"

namespace pg {

namespace {

using namespace umma;

constexpr int BM = 128;
constexpr int LOAD_WARP = 0, TMA_WARP = 1, MMA_WARP = 2, EPI_WARP0 = 4, EPI_WARPS = 8;
constexpr int NTHREADS = (EPI_WARP0 + EPI_WARPS) * 32;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int GROUP_PLANES = 16;   // 128 channels per activation-ring slot

struct PlaneParams {
  const __half* x; int L;              // input planes [B][Cin/8][L][8]
  const int* lens; int in_mask;
  int B, Cin, K, dil, pad;
  const float* bias; const float* bbias; int bbias_ld;
  const uint32_t* tapmask;
  int N, NT, n_ntiles, n_row_tiles, total_tiles;
  int Cout_real, row_mul, L_out;       // output planes [B][Cout_real/8][L_out][8]
  const __half* res16; float res_inv; const float* res32;
  const __half* accin16; const float* accin32;
  __half* out16; float out16_slope; float* out32;
  float out_scale;
  int plane_bytes, KC, PG, n_groups, a_slots, a_slot_bytes, stages, stage_bytes, resident, tmem_cols;
  uint32_t idesc;
  int debug;   // PG_PLANES_DEBUG bitmask (timing experiments only): 1 no A copies, 2 no epilogue I/O, 4 no MMA
};

struct TileCoord { int rt, b, ntile; };
__device__ __forceinline__ TileCoord decode_tile(int id, const PlaneParams& p) {
  TileCoord c;
  c.ntile = id % p.n_ntiles;          // the Cout tiles of one window run back to back (L2 reuse of A)
  const int tmp = id / p.n_ntiles;
  c.rt = tmp % p.n_row_tiles;
  c.b = tmp / p.n_row_tiles;
  return c;
}

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

template <int MT>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_planes_kernel(const __grid_constant__ CUtensorMap wmap, const PlaneParams p) {
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

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < p.a_slots; ++s) {
      mbar_init(&a_full[s], 1);
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

  const int chunks_per_group = p.PG * 8 / p.KC;
  const int n_chunks = p.Cin / p.KC;
  const int planes_total = p.Cin / 8;

  if (warp == LOAD_WARP) {
    // ===== A loader: one bulk copy per plane of the (tile, group) window =====
    const int rows = BM * MT + (p.K - 1) * p.dil;
    uint32_t a_cnt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(tile, p);
      const int tstart = tc.rt * (BM * MT) - p.pad;        // time index of window row 0
      const int len = p.lens ? p.lens[tc.b] : p.L;
      const int t_hi = p.in_mask ? min(p.L, len) : p.L;
      const int lo = min(max(-tstart, 0), rows);           // rows [lo, hi) exist, the rest are zero
      const int hi = min(max(t_hi - tstart, lo), rows);
      const int nz = lo + (rows - hi);
      const __half* xb = p.x + (size_t)tc.b * planes_total * p.L * 8;
      for (int g = 0; g < p.n_groups; ++g, ++a_cnt) {
        const uint32_t slot = a_cnt % (uint32_t)p.a_slots;
        mbar_wait(&a_empty[slot], ((a_cnt / (uint32_t)p.a_slots) & 1u) ^ 1u);
        uint8_t* dst = a_ring + (size_t)slot * p.a_slot_bytes;
        const int pl0 = g * p.PG;
        const int npl = min(p.PG, planes_total - pl0);
        if (nz > 0) {
          for (int i = lane; i < npl * nz; i += 32) {
            const int pl = i / nz, z = i - pl * nz;
            const int r = z < lo ? z : hi + (z - lo);
            *reinterpret_cast<uint4*>(dst + (size_t)pl * p.plane_bytes + (size_t)r * 16) = make_uint4(0u, 0u, 0u, 0u);
          }
          fence_proxy_async_smem();   // generic-proxy zeros -> visible to the tensor core
        }
        __syncwarp();
        const uint32_t bytes = (p.debug & 1) ? 0u : (uint32_t)(hi - lo) * 16u;
        if (lane == 0) mbar_expect_tx(&a_full[slot], bytes * (uint32_t)npl);
        __syncwarp();
        if (bytes && lane < npl)
          bulk_g2s(dst + (size_t)lane * p.plane_bytes + (size_t)lo * 16,
                   xb + ((size_t)(pl0 + lane) * p.L + (tstart + lo)) * 8, bytes, &a_full[slot]);
      }
    }
  } else if (warp == TMA_WARP) {
    // ===== weight producer: (tile, group, chunk, tap) order =====
    if (p.resident) {
      if (elect_one()) {
        mbar_expect_tx(w_ready, (uint32_t)(n_chunks * p.K * p.stage_bytes));
        for (int chunk = 0; chunk < n_chunks; ++chunk)
          for (int tap = 0; tap < p.K; ++tap)
            tma_load_2d(w_ring + (size_t)(chunk * p.K + tap) * p.stage_bytes, &wmap, w_ready, chunk * p.KC,
                        tap * p.N);
      }
      __syncwarp();
    } else {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(tile, p);
        const uint32_t tapmask = p.tapmask ? p.tapmask[tc.ntile] : 0xFFFFFFFFu;
        for (int chunk = 0; chunk < n_chunks; ++chunk) {
          for (int tap = 0; tap < p.K; ++tap) {
            if (!((tapmask >> tap) & 1u)) continue;
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&full[stage], (uint32_t)p.stage_bytes);
              tma_load_2d(w_ring + (size_t)stage * p.stage_bytes, &wmap, &full[stage], chunk * p.KC,
                          tap * p.N + tc.ntile * p.NT);
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
    }
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t_cnt) {
      const TileCoord tc = decode_tile(tile, p);
      const uint32_t tapmask = p.tapmask ? p.tapmask[tc.ntile] : 0xFFFFFFFFu;
      const uint32_t accb = t_cnt & 1u;
      mbar_wait(&acc_empty[accb], ((t_cnt >> 1) & 1u) ^ 1u);
      tcgen05_fence_after();
      const uint32_t d_base = tmem_base + accb * (uint32_t)MT * nt;
      uint32_t started = 0;
      for (int g = 0; g < p.n_groups; ++g, ++a_cnt) {
        const uint32_t slot = a_cnt % (uint32_t)p.a_slots;
        mbar_wait(&a_full[slot], (a_cnt / (uint32_t)p.a_slots) & 1u);
        tcgen05_fence_after();
        const uint32_t a_units = smem_u32(a_ring + (size_t)slot * p.a_slot_bytes) >> 4;
        const int c_begin = g * chunks_per_group;
        const int c_end = min(n_chunks, c_begin + chunks_per_group);
        for (int chunk = c_begin; chunk < c_end; ++chunk) {
          const int chunk_in_group = chunk - c_begin;
          for (int tap = 0; tap < p.K; ++tap) {
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
                    const uint64_t adesc = adesc0 | (uint64_t)(a_k + (uint32_t)m * BM);
                    umma_f16(d_base + (uint32_t)m * nt, adesc, bdesc, idesc, started | (uint32_t)k16);
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
        }
        if (elect_one()) tcgen05_commit(&a_empty[slot]);
        __syncwarp();
      }
      if (elect_one()) tcgen05_commit(&acc_full[accb]);
      __syncwarp();
    }
  } else if (warp >= EPI_WARP0) {
    // ===== epilogue: quarter q = warp % 4 owns TMEM lanes 32q..32q+31 (one output row per lane);
    // the two warps of a quarter alternate over the (row tile m, 32-column block) items.
    const int quarter = warp & 3, grp = (warp - EPI_WARP0) >> 2;
    const int n_cb = p.NT / 32, items = MT * n_cb;
    const int CP = p.Cout_real / 8;
    const bool f32io = p.res32 || p.accin32;
    uint32_t t_cnt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t_cnt) {
      const TileCoord tc = decode_tile(tile, p);
      const int t0 = tc.rt * (BM * MT);
      const int n0 = tc.ntile * p.NT;
      const uint32_t accb = t_cnt & 1u;
      const float* bias = p.bias ? p.bias + n0 : nullptr;
      const float* bbias = p.bbias ? p.bbias + (size_t)tc.b * p.bbias_ld + n0 : nullptr;
      const size_t plane0 = (size_t)tc.b * CP;

      // output offset (in 8-element units) of chunk j of item `it` for this lane's row
      auto chunk_off = [&](int it, int j, bool* ok) -> size_t {
        const int m = it / n_cb, cb = it - m * n_cb;
        const int q = t0 + m * BM + quarter * 32 + lane;
        *ok = q < p.L;
        const int n = n0 + cb * 32 + j * 8;
        const int r = n / p.Cout_real, co = n - r * p.Cout_real;
        return (plane0 + (co >> 3)) * p.L_out + (size_t)(*ok ? q : 0) * p.row_mul + r;
      };
      uint4 rnext[4];
      auto fetch16 = [&](int it) {
        if (!p.res16 || it >= items || (p.debug & 2)) return;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          bool ok;
          const size_t off = chunk_off(it, j, &ok);
          rnext[j] = ok ? __ldg(reinterpret_cast<const uint4*>(p.res16) + off) : make_uint4(0u, 0u, 0u, 0u);
        }
      };
      fetch16(grp);
      mbar_wait(&acc_full[accb], (t_cnt >> 1) & 1u);
      tcgen05_fence_after();
#pragma unroll 1
      for (int it = grp; it < items; it += 2) {
        const int m = it / n_cb, cb = it - m * n_cb;
        uint4 rcur[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) rcur[j] = rnext[j];
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + accb * (uint32_t)(MT * p.NT) +
                      (uint32_t)(m * p.NT + cb * 32), acc);
        if (!f32io) fetch16(it + 2);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          bool ok;
          const size_t off = chunk_off(it, j, &ok);
          const int c = cb * 32 + j * 8;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[j * 8 + i]);
          if (bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          }
          if (bbias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bbias + c));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bbias + c + 4));
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          }
          if (p.res16) {
            float r[8];
            unpack8(rcur[j], r);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += r[i] < 0.f ? r[i] * p.res_inv : r[i];
          }
          if (p.res32 && ok && !(p.debug & 2)) {
            const float4 r0 = __ldg(reinterpret_cast<const float4*>(p.res32) + off * 2);
            const float4 r1 = __ldg(reinterpret_cast<const float4*>(p.res32) + off * 2 + 1);
            v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
            v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
          }
          if (p.out_scale != 1.f) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= p.out_scale;
          }
          if (ok && !(p.debug & 2)) {
            if (p.accin16) {
              float o[8];
              unpack8(*(reinterpret_cast<const uint4*>(p.accin16) + off), o);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += o[i];
            }
            if (p.accin32) {
              const float4 o0 = *(reinterpret_cast<const float4*>(p.accin32) + off * 2);
              const float4 o1 = *(reinterpret_cast<const float4*>(p.accin32) + off * 2 + 1);
              v[0] += o0.x; v[1] += o0.y; v[2] += o0.z; v[3] += o0.w;
              v[4] += o1.x; v[5] += o1.y; v[6] += o1.z; v[7] += o1.w;
            }
            if (p.out32) {
              float4* o = reinterpret_cast<float4*>(p.out32) + off * 2;
              o[0] = make_float4(v[0], v[1], v[2], v[3]);
              o[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
            if (p.out16) {
              if (p.out16_slope != 1.f) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * p.out16_slope;
              }
              *(reinterpret_cast<uint4*>(p.out16) + off) = pack8(v);
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&acc_empty[accb]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tcgen05_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

struct Plan {
  int MT, NT, KC, PG, n_groups, a_slots, a_slot_bytes, stages, stage_bytes, plane_bytes, tmem_cols, resident;
  size_t smem;
};

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

int pick_nt(int n) {
  for (int nt : {256, 128, 64, 32})
    if (n % nt == 0) return nt;
  return 0;
}

bool make_plan(const PlaneConvArgs& a, Plan* out) {
  const int NT = pick_nt(a.N);
  if (!NT) return false;
  const int KC = a.Cin % 64 == 0 ? 64 : 32;
  const int planes = a.Cin / 8;
  int PG = planes < GROUP_PLANES ? planes : GROUP_PLANES;
  if ((PG * 8) % KC) PG = planes;
  if ((PG * 8) % KC) return false;
  const int n_groups = (planes + PG - 1) / PG;
  const int stage_bytes = NT * KC * 2;
  const int n_chunks = a.Cin / KC;
  const size_t w_all = (size_t)n_chunks * a.K * stage_bytes;
  static const int forced_mt = env_int("PG_PLANES_MT", 0), forced_res = env_int("PG_PLANES_RESIDENT", -1),
                   forced_stages = env_int("PG_PLANES_STAGES", 0), forced_slots = env_int("PG_PLANES_SLOTS", 0);
  const bool can_reside = a.N == NT && !a.tapmask && forced_res != 0;
  const size_t fixed = 1024 + 1024;          // alignment slack + barriers
  const size_t budget = (size_t)227 * 1024;
  const long rows128 = (a.L + BM - 1) / BM;
  const int sms = device_sm_count();
  for (int pass = 0; pass < 2; ++pass) {
    const bool resident = pass == 0;
    if (resident && !can_reside) continue;
    if (!resident && forced_res == 1 && can_reside) continue;
    for (int MT : {4, 2, 1}) {
      if (forced_mt > 0 && MT != forced_mt) continue;
      if (2 * MT * NT > 512) continue;
      // keep at least ~2 tiles per SM before growing the tile
      if (forced_mt == 0 && MT > 1 && ((rows128 + MT - 1) / MT) * a.B * (a.N / NT) < 2L * sms) continue;
      const int rows = BM * MT + (a.K - 1) * a.dil;
      const int plane_bytes = 16 * ((rows + 7) & ~7);
      const int a_slot_bytes = (PG * plane_bytes + 127) & ~127;
      const size_t w_min = resident ? w_all : 2 * (size_t)stage_bytes;
      if (fixed + 2 * (size_t)a_slot_bytes + w_min > budget) continue;
      size_t left = budget - fixed - 2 * (size_t)a_slot_bytes - w_min;
      int a_slots = 2, stages = resident ? n_chunks * a.K : 2;
      if (!resident) {
        const int max_stages = forced_stages > 0 ? forced_stages : 6;
        while (stages < max_stages && stages < n_chunks * a.K && left >= (size_t)stage_bytes) {
          ++stages;
          left -= stage_bytes;
        }
      }
      const int max_slots = forced_slots > 0 ? forced_slots : 4;
      while (a_slots < max_slots && left >= (size_t)a_slot_bytes) {
        ++a_slots;
        left -= a_slot_bytes;
      }
      int tm = 32;
      while (tm < 2 * MT * NT) tm <<= 1;
      *out = Plan{MT, NT, KC, PG, n_groups, a_slots, a_slot_bytes, stages, stage_bytes, plane_bytes, tm,
                  resident ? 1 : 0, fixed + (size_t)a_slots * a_slot_bytes + (size_t)stages * stage_bytes};
      return true;
    }
  }
  return false;
}

template <int MT>
cudaError_t launch_t(const PlaneConvArgs& a, const Plan& pl, cudaStream_t s) {
  CUtensorMap wmap;
  if (!get_weight_map(a.w16, a.Cin, a.N, a.K, pl.KC, pl.NT, &wmap)) return cudaErrorNotSupported;
  PlaneParams p;
  p.x = a.x; p.L = a.L; p.lens = a.lens; p.in_mask = a.in_mask;
  p.B = a.B; p.Cin = a.Cin; p.K = a.K; p.dil = a.dil; p.pad = a.pad;
  p.bias = a.bias; p.bbias = a.bbias; p.bbias_ld = a.bbias_ld; p.tapmask = a.tapmask;
  p.N = a.N; p.NT = pl.NT; p.n_ntiles = a.N / pl.NT;
  p.n_row_tiles = (a.L + BM * MT - 1) / (BM * MT);
  p.total_tiles = p.n_row_tiles * a.B * p.n_ntiles;
  p.Cout_real = a.Cout_real; p.row_mul = a.row_mul; p.L_out = a.L * a.row_mul;
  p.res16 = a.res16; p.res_inv = a.res_inv; p.res32 = a.res32;
  p.accin16 = a.accin16; p.accin32 = a.accin32;
  p.out16 = a.out16; p.out16_slope = a.out16_slope; p.out32 = a.out32; p.out_scale = a.out_scale;
  p.plane_bytes = pl.plane_bytes; p.KC = pl.KC; p.PG = pl.PG; p.n_groups = pl.n_groups;
  p.a_slots = pl.a_slots; p.a_slot_bytes = pl.a_slot_bytes; p.stages = pl.stages;
  p.stage_bytes = pl.stage_bytes; p.resident = pl.resident; p.tmem_cols = pl.tmem_cols;
  p.idesc = make_idesc(BM, pl.NT);
  static const int dbg = env_int("PG_PLANES_DEBUG", 0);
  p.debug = dbg;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_planes_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  int grid = device_sm_count();
  if (grid > p.total_tiles) grid = p.total_tiles;
  conv_planes_kernel<MT><<<grid, NTHREADS, pl.smem, s>>>(wmap, p);
  return cudaGetLastError();
}

}  // namespace

int plane_pick_nt(int n) { return pick_nt(n); }

bool plane_conv_supported(const PlaneConvArgs& a) {
  if (!a.x || !a.w16 || (!a.out16 && !a.out32)) return false;
  if (a.Cin % 32 || a.Cin < 32 || a.N % 32) return false;
  if (a.K < 1 || a.K > 32 || a.L <= 0 || a.B <= 0) return false;
  if (a.row_mul < 1 || a.Cout_real % 8 || a.Cout_real * a.row_mul != a.N) return false;
  if (a.in_mask && !a.lens) return false;
  Plan pl;
  return make_plan(a, &pl);
}

cudaError_t launch_conv_planes(const PlaneConvArgs& a, cudaStream_t s) {
  Plan pl;
  if (!plane_conv_supported(a) || !make_plan(a, &pl)) return cudaErrorInvalidValue;
  switch (pl.MT) {
    case 1: return launch_t<1>(a, pl, s);
    case 2: return launch_t<2>(a, pl, s);
    case 4: return launch_t<4>(a, pl, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace pg
