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
//   warp 3      residual loader: bulk copies of the epilogue's residual items into a shared ring
//   warps 4-19  epilogue: TMEM -> registers -> bias / residual / scale / accumulate / leaky-ReLU
//               -> global.  Two accumulator sets so the epilogue of tile i overlaps the MMAs of i+1.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "pg_common.cuh"
#include "pg_umma.cuh"

namespace pg {

namespace {

using namespace umma;

constexpr int BM = 128;
constexpr int LOAD_WARP = 0, TMA_WARP = 1, MMA_WARP = 2, RES_WARP = 3, EPI_WARP0 = 4, EPI_WARPS = 16;
constexpr int EPI_GROUPS = EPI_WARPS / 4;   // warps per TMEM lane quarter
constexpr int EPI_COLS = 16;                // accumulator columns per epilogue item
// epilogue specialisations: the two shapes that make up 5/6 of the ResBlock convs get straight-line code
constexpr int EPI_GENERIC = 0, EPI_C1 = 1 /* y16 = lrelu(acc + bias) */, EPI_C2 = 2 /* y16 = lrelu(acc + bias + raw(res16)) */,
              EPI_C3 = 3 /* last conv2 of a ResBlock: y16 = lrelu((acc + bias + raw(res16)) * scale [+ accin16]) */,
              EPI_UPS = 4 /* polyphase upsampler / plain conv over several Cout tiles: y16 = lrelu(acc + bias [+ accin16]) at row t*row_mul + phase */;
constexpr int NTHREADS = (EPI_WARP0 + EPI_WARPS) * 32;
constexpr int GROUP_PLANES = 16;   // 128 channels per activation-ring slot
constexpr int NZ_MAXK = 8;         // longest noise conv fused into an upsampler epilogue (taps)

struct PlaneParams {
  const __half* x; int L;              // input planes [B][Cin/8][L][8]
  const int* lens; int in_mask;
  const int* tlen; int len_mul;        // hard end of row b (tile rows): tlen[b]*len_mul, nullptr = L
  int B, Cin, K, dil, pad;
  const float* bias; const float* bbias; int bbias_ld;
  const uint32_t* tapmask;
  int N, NT, n_ntiles, n_row_tiles, total_tiles;
  int Cout_real, row_mul, L_out;       // output planes [B][Cout_real/8][L_out][8]
  const __half* res16; float res_inv; const float* res32;
  const __half* accin16; const float* accin32;
  __half* out16; float out16_slope; float* out32;
  float out_scale;
  // fused NSF source injection (nsf.py:131): + bn[co] + sum_j wn[j][co] * src[b][t_out*stride + j - pad]
  const float* nz_src; const float* nz_w; const float* nz_b; int nz_k, nz_stride, nz_pad, nz_len;
  __half* out_lo;                      // hi/lo output: out16 = f16(lrelu(v)), out_lo = its remainder
  int plane_bytes, KC, PG, n_groups, a_slots, a_slot_bytes, stages, stage_bytes, resident, tmem_cols;
  int r_slots, r_slot_bytes, res_cols;  // residual ring: one slot = 128 rows x res_cols columns
  int bias_smem;                        // bias[N] staged in shared memory (N <= 1024)
  uint32_t idesc;
  int nz_prefetch;                      // EPI_UPS: fetch the noise conv's source samples one item ahead (PG_NZ_PREFETCH, A/B aid)
  int debug;   // PG_PLANES_DEBUG bitmask (timing experiments only): 1 no A copies, 2 no epilogue I/O, 4 no MMA,
               // 8 trace, 16 consumers skip their waits (free-running roles), 32 epilogue idle, 64 no producers
};

// PG_PLANES_DEBUG & 8: CTA 0 records (role, event, tile, globaltimer) -- timing experiments only
__device__ unsigned long long g_ptrace[4096];
__device__ unsigned int g_ptrace_n;
__device__ __forceinline__ void trace(int dbg, int role, int ev, int tile) {
  if ((dbg & 8) && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned int i = atomicAdd(&g_ptrace_n, 1u);
    if (i < 4096) g_ptrace[i] = (t << 20) | ((unsigned long long)role << 16) | ((unsigned long long)ev << 12) | (unsigned)(tile & 0xFFF);
  }
}

struct TileCoord { int rt, b, ntile; };
__device__ __forceinline__ TileCoord decode_tile(int id, const PlaneParams& p) {
  TileCoord c;
  c.ntile = id % p.n_ntiles;          // the Cout tiles of one window run back to back (L2 reuse of A)
  const int tmp = id / p.n_ntiles;
  c.rt = tmp % p.n_row_tiles;
  c.b = tmp / p.n_row_tiles;
  return c;
}

// hard end (in tile rows) of batch row b; tiles that start at or beyond it are skipped by EVERY role
// (same test everywhere, so the per-role ring / accumulator counters stay in step)
__device__ __forceinline__ int row_end(const PlaneParams& p, int b) {
  return p.tlen ? min(p.L, p.tlen[b] * p.len_mul) : p.L;
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

// SWAP (C = 128 ResBlock shapes, MT == 2, NT == 128): the operands change roles -- the weight tile is
// the M = 128 operand (rows = Cout) and the MT*128 = 256 time rows of the window are the N operand,
// D[Cout][time].  One N = 256 MMA then reads 16 KB (weights) + 32 KB (window) of shared memory per 512
// clk = 96 B/clk instead of the 128 B/clk of two N = 128 MMAs, which is what caps these layers at
// ~60 % tensor-pipe duty.  TMEM lanes become channels, columns time, so the epilogue
// transposes 8-lane x 8-(row pair) blocks with warp shuffles back into 16-byte plane rows.
template <int MT, int KC16, int EPI, bool DBG, bool SWAP = false>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_planes_kernel(const __grid_constant__ CUtensorMap wmap, const PlaneParams p) {
  pdl_launch_dependents();     // the next kernel of the stream may launch and run its prologue (pg_common.cuh)
  const int dbg = DBG ? p.debug : 0;   // production instantiation: every debug switch folds away
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_ring = smem;
  uint8_t* a_ring = w_ring + (size_t)p.stages * p.stage_bytes;
  uint8_t* r_ring = a_ring + (size_t)p.a_slots * p.a_slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(r_ring + (size_t)p.r_slots * p.r_slot_bytes);
  uint64_t* full = bars;                      // [stages]
  uint64_t* empty = full + p.stages;          // [stages]
  uint64_t* a_full = empty + p.stages;        // [a_slots]
  uint64_t* a_empty = a_full + p.a_slots;     // [a_slots]
  uint64_t* acc_full = a_empty + p.a_slots;   // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint64_t* w_ready = acc_empty + 2;          // resident weights landed
  uint64_t* r_full = w_ready + 1;             // [r_slots]
  uint64_t* r_empty = r_full + p.r_slots;     // [r_slots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_empty + p.r_slots);
  float* bias_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));   // [N] when p.bias_smem

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
      mbar_init(&acc_empty[s], EPI_WARPS);
    }
    for (int s = 0; s < p.r_slots; ++s) {
      mbar_init(&r_full[s], 1);
      mbar_init(&r_empty[s], (uint32_t)(4 * p.res_cols / EPI_COLS));   // every warp-item that reads the slot
    }
    mbar_init(w_ready, 1);
    fence_barrier_init();
  }
  if (p.bias_smem)
    for (int i = tid; i < p.N; i += NTHREADS) bias_s[i] = p.bias[i];
  // fused source injection: the noise conv's bias and [k][C] weights sit right behind the bias vector
  float* nz_s = bias_s + (p.bias_smem ? ((p.N + 3) & ~3) : 0);
  if (p.nz_src)
    for (int i = tid; i < (p.nz_k + 1) * p.Cout_real; i += NTHREADS)
      nz_s[i] = i < p.Cout_real ? p.nz_b[i] : p.nz_w[i - p.Cout_real];
  if (warp == MMA_WARP) tcgen05_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  if (warp == TMA_WARP && lane == 0)
    asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) trace(dbg, 4, 0, 0);
  // activations in global memory belong to the predecessor until it has completed; the weight producer and the MMA
  // issuer touch constants / shared memory / TMEM only and run ahead
  if (warp != TMA_WARP && warp != MMA_WARP) pdl_wait();

  const int chunks_per_group = p.PG * 8 / p.KC;
  const int n_chunks = p.Cin / p.KC;
  const int planes_total = p.Cin / 8;

  const bool freerun = (dbg & 16) != 0;
  if (warp == LOAD_WARP && !(dbg & 64)) {
    // ===== A loader: one bulk copy per plane of the (tile, group) window =====
    const int rows = BM * MT + (p.K - 1) * p.dil;
    uint32_t a_cnt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(tile, p);
      const int t_end = row_end(p, tc.b);
      if (tc.rt * (BM * MT) >= t_end) continue;            // dead tile (past the row's hard end)
      const int tstart = tc.rt * (BM * MT) - p.pad;        // time index of window row 0
      const int len = p.lens ? p.lens[tc.b] : p.L;
      const int t_hi = p.in_mask ? min(t_end, len) : t_end;
      const int lo = min(max(-tstart, 0), rows);           // rows [lo, hi) exist, the rest are zero
      const int hi = min(max(t_hi - tstart, lo), rows);
      const int nz = lo + (rows - hi);
      const __half* xb = p.x + (size_t)tc.b * planes_total * p.L * 8;
      for (int g = 0; g < p.n_groups; ++g, ++a_cnt) {
        const uint32_t slot = a_cnt % (uint32_t)p.a_slots;
        mbar_wait(&a_empty[slot], ((a_cnt / (uint32_t)p.a_slots) & 1u) ^ 1u);
        if (lane == 0) trace(dbg, 0, 1, (int)a_cnt);
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
        const uint32_t bytes = (dbg & 1) ? 0u : (uint32_t)(hi - lo) * 16u;
        if (lane == 0) mbar_expect_tx(&a_full[slot], bytes * (uint32_t)npl);
        __syncwarp();
        if (bytes && lane < npl)
          bulk_g2s(dst + (size_t)lane * p.plane_bytes + (size_t)lo * 16,
                   xb + ((size_t)(pl0 + lane) * p.L + (tstart + lo)) * 8, bytes, &a_full[slot]);
      }
    }
  } else if (warp == TMA_WARP && !(dbg & 64)) {
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
        if (tc.rt * (BM * MT) >= row_end(p, tc.b)) continue;
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
  } else if (warp == MMA_WARP && !(dbg & 128)) {
    // ===== MMA issuer: ONE elected lane runs the whole role (waits, issues, commits); inside the
    // single-lane region the compiler moves operands to uniform registers without waterfall loops =====
    if (elect_one()) {
    const uint32_t w_layout = p.KC == 64 ? LAYOUT_SW128 : LAYOUT_SW64;
    const uint32_t w_sbo = 8u * (uint32_t)p.KC * 2u;
    const uint64_t adesc0 = make_desc(0u, (uint32_t)p.plane_bytes, 128u, LAYOUT_NONE);
    const uint64_t bdesc0 = make_desc(0u, 0u, w_sbo, w_layout);
    const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
    const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;   // LBO field, start = 0
    const uint32_t plane_units = (uint32_t)p.plane_bytes >> 4;
    const uint32_t plane2 = 2u * plane_units;
    const int kc8 = p.KC / 8;
    const uint32_t nt = (uint32_t)p.NT, idesc = p.idesc;
    const bool do_mma = !(dbg & 4), do_commit = !(dbg & 512);
    const uint32_t w_ring_units = smem_u32(w_ring) >> 4, stage_units = (uint32_t)p.stage_bytes >> 4;
    const uint32_t a_ring_units = smem_u32(a_ring) >> 4, slot_units = (uint32_t)p.a_slot_bytes >> 4;
    int stage = 0;
    uint32_t phase = 0, a_cnt = 0, t_cnt = 0;
    if (p.resident && !freerun) {
      mbar_wait(w_ready, 0);
      tcgen05_fence_after();
    }
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      {
        const TileCoord tc = decode_tile(tile, p);
        if (tc.rt * (BM * MT) >= row_end(p, tc.b)) continue;
      }
      const uint32_t tapmask = p.tapmask ? p.tapmask[(tile % p.n_ntiles)] : 0xFFFFFFFFu;
      const uint32_t accb = t_cnt & 1u;
      if (!freerun) mbar_wait(&acc_empty[accb], ((t_cnt >> 1) & 1u) ^ 1u);
      tcgen05_fence_after();
      trace(dbg, 1, 1, (int)t_cnt);
      const uint32_t d_base = tmem_base + accb * (uint32_t)MT * nt;
      uint32_t started = 0;
      for (int g = 0; g < p.n_groups; ++g, ++a_cnt) {
        const uint32_t slot = a_cnt % (uint32_t)p.a_slots;
        if (!freerun) mbar_wait(&a_full[slot], (a_cnt / (uint32_t)p.a_slots) & 1u);
        tcgen05_fence_after();
        trace(dbg, 1, 2, (int)t_cnt);
        const uint32_t a_units = a_ring_units + slot * slot_units;
        const int c_begin = g * chunks_per_group;
        const int c_end = min(n_chunks, c_begin + chunks_per_group);
        for (int chunk = c_begin; chunk < c_end; ++chunk) {
          const uint32_t a_chunk = a_lo0 + a_units + (uint32_t)((chunk - c_begin) * kc8) * plane_units;
          uint32_t w_res = b_lo0 + w_ring_units + (uint32_t)(chunk * p.K) * stage_units;   // resident tile
#pragma unroll 1
          for (int tap = 0; tap < p.K; ++tap, w_res += stage_units) {
            if (!((tapmask >> tap) & 1u)) continue;
            uint32_t w_lo = w_res;
            if (!p.resident) {
              if (!freerun) mbar_wait(&full[stage], phase);
              tcgen05_fence_after();
              w_lo = b_lo0 + w_ring_units + (uint32_t)stage * stage_units;
            }
            const uint32_t a_lo = a_chunk + (uint32_t)(tap * p.dil);
#pragma unroll
            for (int k16 = 0; k16 < KC16; ++k16) {
              if (SWAP) {
                if (do_mma)   // D[Cout 128][time MT*128] += W[Cout][16] * X[time][16]^T
                  umma_f16_lh(d_base, w_lo + 2u * (uint32_t)k16, b_hi, a_lo + (uint32_t)k16 * plane2, a_hi, idesc,
                              started | (uint32_t)k16);
              } else {
#pragma unroll
                for (int m = 0; m < MT; ++m)
                  if (do_mma)
                    umma_f16_lh(d_base + (uint32_t)m * nt, a_lo + (uint32_t)k16 * plane2 + (uint32_t)(m * BM), a_hi,
                                w_lo + 2u * (uint32_t)k16, b_hi, idesc, started | (uint32_t)k16);
              }
            }
            started = 1;
            if (!p.resident) {
              if (do_commit) tcgen05_commit(&empty[stage]);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
        if (do_commit) tcgen05_commit(&a_empty[slot]);
      }
      if (do_commit) tcgen05_commit(&acc_full[accb]);
      trace(dbg, 1, 3, (int)t_cnt);
      ++t_cnt;
    }
    }
  } else if (warp == RES_WARP && p.r_slots > 0 && !(dbg & (64 | 16384))) {
    // ===== residual loader: slot = (row tile m, res_cols-column block) -> res_cols/8 planes x 128 rows.
    // The <= 4 slots of a tile are issued together: lanes [8w, 8w + res_cols/8) serve slot w.
    const int n_rb = p.NT / p.res_cols, per_tile = MT * n_rb, rb_shift = 31 - __clz(n_rb);
    const int rplanes = p.res_cols / 8;
    const int CP = p.Cout_real / 8;
    const uint32_t esz = p.res32 ? 32u : 16u;   // bytes per (row, 8-channel chunk)
    const char* rbase = p.res32 ? reinterpret_cast<const char*>(p.res32) : reinterpret_cast<const char*>(p.res16);
    const int sub = lane >> 3, pl = lane & 7;
    const int batch = min(min(per_tile, p.r_slots), 4);   // slots issued together (distinct ring entries)
    uint32_t s_cnt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(tile, p);
      const int t0 = tc.rt * (BM * MT), n0 = tc.ntile * p.NT;
      if (t0 >= row_end(p, tc.b)) continue;
      const size_t plane0 = (size_t)tc.b * CP;
      for (int base = 0; base < per_tile; base += batch) {
        const int which = base + sub;
        const bool mine = sub < batch && which < per_tile && pl < rplanes;
        const uint32_t sidx = s_cnt + (uint32_t)which;
        const uint32_t slot = sidx % (uint32_t)p.r_slots;
        const int m = which >> rb_shift, rb = which & (n_rb - 1);
        const int row0 = t0 + m * BM;
        const int nrows = min(BM, p.L - row0);
        const uint32_t bytes = (nrows > 0 && !(dbg & (2 | 4096))) ? (uint32_t)nrows * esz : 0u;
        if (mine) mbar_wait(&r_empty[slot], ((sidx / (uint32_t)p.r_slots) & 1u) ^ 1u);
        __syncwarp();
        if (mine && pl == 0) mbar_expect_tx(&r_full[slot], (uint32_t)rplanes * bytes);
        __syncwarp();
        if (mine && bytes) {
          const int co = n0 + rb * p.res_cols + pl * 8;      // residual: row_mul == 1, column == channel
          const size_t off = (plane0 + (co >> 3)) * p.L_out + row0;
          bulk_g2s(r_ring + (size_t)slot * p.r_slot_bytes + (size_t)pl * (BM * esz), rbase + off * esz, bytes,
                   &r_full[slot]);
        }
      }
      s_cnt += (uint32_t)per_tile;
    }
  } else if (warp >= EPI_WARP0 && EPI != EPI_GENERIC && SWAP) {
    // ===== SWAP epilogue (EPI_C1 / EPI_C2, NT == Cout == 128): TMEM lane = channel, column = time row.
    // Item = 16 time columns: tcgen05.ld -> + bias (per lane) -> 8x8 shuffle transpose inside each group of
    // 8 lanes (= one 8-channel plane) -> lane i of the group owns rows 2i, 2i+1 of the block as two 16-byte
    // plane rows -> (+ residual) -> leaky-ReLU -> store.
    const int quarter = warp & 3, grp = (warp - EPI_WARP0) >> 2;
    constexpr int ITEMS = MT * BM / EPI_COLS;
    const int CP = p.Cout_real / 8;
    const float slope = p.out16_slope, rinv = p.res_inv;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int j8 = lane >> 3, i8 = lane & 7;
    const int plane = quarter * 4 + j8;
    const float bias_c = bias_s[quarter * 32 + lane];
    uint32_t t_cnt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int rt = tile % p.n_row_tiles, b = tile / p.n_row_tiles;     // n_ntiles == 1
      if (rt * (BM * MT) >= row_end(p, b)) continue;
      const uint32_t accb = t_cnt & 1u;
      uint4* out_b = reinterpret_cast<uint4*>(p.out16) + ((size_t)b * CP + plane) * p.L;
      mbar_wait(&acc_full[accb], (t_cnt >> 1) & 1u);
      tcgen05_fence_after();
#pragma unroll 1
      for (int it = grp; it < ITEMS; it += EPI_GROUPS) {
        const int col0 = it * EPI_COLS;                  // time column inside the tile
        const int m = col0 / BM;
        const int row = rt * (BM * MT) + col0 + 2 * i8;  // first of this lane's two output rows
        uint4 rcur[2];
        int ring_slot = 0;
        if (EPI == EPI_C2) {
          // ring slot = (row tile m, 64-channel block): the four planes of a quarter share one slot
          const uint32_t sidx = (t_cnt * (uint32_t)MT + (uint32_t)m) * (uint32_t)(p.NT / p.res_cols) +
                                (uint32_t)((quarter * 32) / p.res_cols);
          const uint32_t slot = sidx % (uint32_t)p.r_slots;
          mbar_wait(&r_full[slot], (sidx / (uint32_t)p.r_slots) & 1u);
          const uint8_t* rs = r_ring + (size_t)slot * p.r_slot_bytes +
                              (size_t)(((quarter * 32) % p.res_cols) / 8 + j8) * (BM * 16) +
                              (size_t)((col0 % BM) + 2 * i8) * 16;
          rcur[0] = *reinterpret_cast<const uint4*>(rs);
          rcur[1] = *reinterpret_cast<const uint4*>(rs + 16);
          ring_slot = (int)slot;
        }
        uint32_t acc[EPI_COLS];
        tmem_ld16(lane_taddr + accb * (uint32_t)(MT * BM) + (uint32_t)col0, acc);
        if (EPI == EPI_C2) {   // release the slot only after the loads have landed (see mbar_arrive_dep)
          uint32_t dep = rcur[0].x ^ rcur[0].w ^ rcur[1].x ^ rcur[1].w;
          asm volatile("" : "+r"(dep));
          __syncwarp();
          if (lane == 0) mbar_arrive_dep(&r_empty[ring_slot], dep);
        }
        float a[EPI_COLS];
#pragma unroll
        for (int i = 0; i < EPI_COLS; ++i) a[i] = __uint_as_float(acc[i]) + bias_c;
        // transpose: element e = (a[2e], a[2e+1]) = time pair e of this lane's channel; afterwards
        // element k = time pair i8 of channel k of the plane
#pragma unroll
        for (int mask = 4; mask >= 1; mask >>= 1) {
          const bool up = (i8 & mask) != 0;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (e & mask) continue;
            const int f = e | mask;
            const float s0 = up ? a[2 * e] : a[2 * f], s1 = up ? a[2 * e + 1] : a[2 * f + 1];
            const float r0 = __shfl_xor_sync(0xffffffffu, s0, mask), r1 = __shfl_xor_sync(0xffffffffu, s1, mask);
            if (up) {
              a[2 * e] = r0;
              a[2 * e + 1] = r1;
            } else {
              a[2 * f] = r0;
              a[2 * f + 1] = r1;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = a[2 * k + j];
          if (EPI == EPI_C2) {
            float rr[8];
            unpack8(rcur[j], rr);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += fminf(rr[k], rr[k] * rinv);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], v[k] * slope);
          if (row + j < p.L) out_b[row + j] = pack8(v);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[accb]);
      ++t_cnt;
    }
  } else if (warp >= EPI_WARP0 && EPI == EPI_UPS) {
    // ===== straight-line epilogue of the polyphase upsamplers (and of any bias-only conv with several Cout tiles):
    // per item tcgen05.ld 16 columns -> + bias (shared memory) [+ accin16] -> leaky-ReLU -> 2 x 16 B stores at output
    // row t*row_mul + phase.  The generic epilogue spent ~340 instructions per item on run-time switches and fetched
    // the bias of a > 1024-column layer through __ldg (ncu: 6.8 warps stalled on long scoreboard per issue).
    const int quarter = warp & 3, grp = (warp - EPI_WARP0) >> 2;
    const int n_cb = p.NT / EPI_COLS, items = MT * n_cb, cb_shift = 31 - __clz(n_cb);
    const int CP = p.Cout_real / 8;
    const int cout_shift = 31 - __clz(p.Cout_real);
    const float slope = p.out16_slope;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t t_cnt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(tile, p);
      const int t0 = tc.rt * (BM * MT);
      if (t0 >= row_end(p, tc.b)) continue;
      const int n0 = tc.ntile * p.NT;
      const uint32_t accb = t_cnt & 1u;
      const size_t plane0 = (size_t)tc.b * CP;
      const int qbase = t0 + quarter * 32 + lane;
      // source samples under the noise conv's taps of item `it` (this lane's output row); fetched one item ahead --
      // the first one before the accumulator wait -- so their L2 latency hides behind the previous item
      const float* nz_sp = p.nz_src ? p.nz_src + (size_t)tc.b * p.nz_len : nullptr;
      auto load_sv = [&](int it, float (&sv)[NZ_MAXK]) {
        const int q = qbase + (it >> cb_shift) * BM;
        const int r = (n0 + (it & (n_cb - 1)) * EPI_COLS) >> cout_shift;
        const int s0 = (q * p.row_mul + r) * p.nz_stride - p.nz_pad;
#pragma unroll
        for (int jj = 0; jj < NZ_MAXK; ++jj) {
          if (jj >= p.nz_k) break;       // uniform early exit: a 1-tap stage issues one load, not eight predicated ones
          const int sn = s0 + jj;
          sv[jj] = (q < p.L && sn >= 0 && sn < p.nz_len) ? __ldg(nz_sp + sn) : 0.f;
        }
      };
      float sv_next[NZ_MAXK];
      if (nz_sp && p.nz_prefetch && grp < items) load_sv(grp, sv_next);
      mbar_wait(&acc_full[accb], (t_cnt >> 1) & 1u);
      tcgen05_fence_after();
#pragma unroll 1
      for (int it = grp; it < items; it += EPI_GROUPS) {
        const int m = it >> cb_shift, cb = it & (n_cb - 1);
        const int q = qbase + m * BM;
        const int c0 = cb * EPI_COLS;
        const int n = n0 + c0;                             // GEMM column -> (phase r, channel co)
        const int r = n >> cout_shift, co = n & (p.Cout_real - 1);
        const size_t off0 = (plane0 + (co >> 3)) * p.L_out + (size_t)(q < p.L ? q : 0) * p.row_mul + r;
        float sv[NZ_MAXK];
        if (nz_sp && p.nz_prefetch) {
#pragma unroll
          for (int jj = 0; jj < NZ_MAXK; ++jj) {
            if (jj >= p.nz_k) break;
            sv[jj] = sv_next[jj];
          }
          if (it + EPI_GROUPS < items) load_sv(it + EPI_GROUPS, sv_next);
        } else if (nz_sp) {
          load_sv(it, sv);
        }
        uint4 ain[2];
        if (p.accin16 && q < p.L) {
          ain[0] = *(reinterpret_cast<const uint4*>(p.accin16) + off0);
          ain[1] = *(reinterpret_cast<const uint4*>(p.accin16) + off0 + p.L_out);
        }
        float bv[EPI_COLS];
#pragma unroll
        for (int i = 0; i < EPI_COLS / 4; ++i) {
          const float4 b4 = *(reinterpret_cast<const float4*>(bias_s + n) + i);
          bv[4 * i] = b4.x; bv[4 * i + 1] = b4.y; bv[4 * i + 2] = b4.z; bv[4 * i + 3] = b4.w;
        }
        uint32_t acc[EPI_COLS];
        tmem_ld16(lane_taddr + accb * (uint32_t)(MT * p.NT) + (uint32_t)(m * p.NT + c0), acc);
        if (q < p.L) {
          uint4* dst = reinterpret_cast<uint4*>(p.out16) + off0;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[j * 8 + i]) + bv[j * 8 + i];
            if (p.nz_src) {     // x + noise_convs[i](har_source): [bias | k x C] weights in shared memory
              const float* wb = nz_s + co + j * 8;
              const float4 b0 = *reinterpret_cast<const float4*>(wb), b1 = *reinterpret_cast<const float4*>(wb + 4);
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
#pragma unroll
              for (int jj = 0; jj < NZ_MAXK; ++jj) {
                if (jj >= p.nz_k) break;
                const float4 w0 = *reinterpret_cast<const float4*>(wb + (size_t)(jj + 1) * p.Cout_real);
                const float4 w1 = *reinterpret_cast<const float4*>(wb + (size_t)(jj + 1) * p.Cout_real + 4);
                const float x = sv[jj];
                v[0] = fmaf(w0.x, x, v[0]); v[1] = fmaf(w0.y, x, v[1]);
                v[2] = fmaf(w0.z, x, v[2]); v[3] = fmaf(w0.w, x, v[3]);
                v[4] = fmaf(w1.x, x, v[4]); v[5] = fmaf(w1.y, x, v[5]);
                v[6] = fmaf(w1.z, x, v[6]); v[7] = fmaf(w1.w, x, v[7]);
              }
            }
            if (p.accin16) {
              float av[8];
              unpack8(ain[j], av);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += av[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], v[i] * slope);     // slope <= 1 (1: raw)
            const uint4 hi = pack8(v);
            dst[(size_t)j * p.L_out] = hi;
            if (p.out_lo) {       // hi/lo stream of the last stage: the f16 rounding remainder
              float hv[8];
              unpack8(hi, hv);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] -= hv[i];
              *(reinterpret_cast<uint4*>(p.out_lo) + off0 + (size_t)j * p.L_out) = pack8(v);
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[accb]);
      ++t_cnt;
    }
  } else if (warp >= EPI_WARP0 && EPI != EPI_GENERIC) {
    // ===== specialised epilogue (EPI_C1 / EPI_C2): one Cout tile, row_mul == 1, bias in shared memory,
    // f16 L-form output; per item: tcgen05.ld 16 columns -> + bias (+ residual) -> leaky-ReLU -> 2 x 16 B.
    const int quarter = warp & 3, grp = (warp - EPI_WARP0) >> 2;
    const int n_cb = p.NT / EPI_COLS, items = MT * n_cb, cb_shift = 31 - __clz(n_cb);
    const int CP = p.Cout_real / 8;
    const float slope = p.out16_slope, rinv = p.res_inv, scale = p.out_scale;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int rrow16 = (quarter * 32 + lane) * 16;
    uint32_t t_cnt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int rt = tile % p.n_row_tiles, b = tile / p.n_row_tiles;     // n_ntiles == 1
      if (rt * (BM * MT) >= row_end(p, b)) continue;
      const int qbase = rt * (BM * MT) + quarter * 32 + lane;
      const uint32_t accb = t_cnt & 1u;
      uint4* out_b = reinterpret_cast<uint4*>(p.out16) + (size_t)b * CP * p.L;
      const uint4* acc_b = reinterpret_cast<const uint4*>(p.accin16) + (size_t)b * CP * p.L;
      mbar_wait(&acc_full[accb], (t_cnt >> 1) & 1u);
      tcgen05_fence_after();
#pragma unroll 1
      for (int it = grp; it < items; it += EPI_GROUPS) {
        const int m = it >> cb_shift, cb = it & (n_cb - 1);
        const int q = qbase + m * BM;
        const int c0 = cb * EPI_COLS;
        uint4 ain[2];
        if (EPI == EPI_C3 && p.accin16 && q < p.L) {      // running mean: issued first, used last
          const uint4* g = acc_b + (size_t)(c0 >> 3) * p.L + q;
          ain[0] = g[0];
          ain[1] = g[p.L];
        }
        float bv[EPI_COLS];
#pragma unroll
        for (int i = 0; i < EPI_COLS / 4; ++i) {
          const float4 b4 = *(reinterpret_cast<const float4*>(bias_s + c0) + i);
          bv[4 * i] = b4.x; bv[4 * i + 1] = b4.y; bv[4 * i + 2] = b4.z; bv[4 * i + 3] = b4.w;
        }
        uint4 rcur[2];
        int ring_slot = 0;
        if (EPI == EPI_C2 || EPI == EPI_C3) {
          const uint32_t sidx = (t_cnt * (uint32_t)MT + (uint32_t)m) * (uint32_t)(p.NT / p.res_cols) + (uint32_t)(c0 / p.res_cols);
          const uint32_t slot = sidx % (uint32_t)p.r_slots;
          mbar_wait(&r_full[slot], (sidx / (uint32_t)p.r_slots) & 1u);
          const uint8_t* rs = r_ring + (size_t)slot * p.r_slot_bytes + (size_t)((c0 % p.res_cols) >> 3) * (BM * 16) + rrow16;
          rcur[0] = *reinterpret_cast<const uint4*>(rs);
          rcur[1] = *reinterpret_cast<const uint4*>(rs + BM * 16);
          ring_slot = (int)slot;
        }
        uint32_t acc[EPI_COLS];
        tmem_ld16(lane_taddr + accb * (uint32_t)(MT * p.NT) + (uint32_t)(m * p.NT + c0), acc);
        if (EPI == EPI_C2 || EPI == EPI_C3) {   // release the slot only after the loads have landed (see mbar_arrive_dep)
          uint32_t dep = rcur[0].x ^ rcur[0].w ^ rcur[1].x ^ rcur[1].w;
          asm volatile("" : "+r"(dep));
          __syncwarp();
          if (lane == 0) mbar_arrive_dep(&r_empty[ring_slot], dep);
        }
        if (q < p.L) {
          uint4* dst = out_b + (size_t)(c0 >> 3) * p.L + q;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[j * 8 + i]) + bv[j * 8 + i];
            if (EPI == EPI_C2 || EPI == EPI_C3) {
              float rr[8];
              unpack8(rcur[j], rr);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += fminf(rr[i], rr[i] * rinv);
            }
            if (EPI == EPI_C3) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] *= scale;
              if (p.accin16) {
                float av[8];
                unpack8(ain[j], av);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] += av[i];
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], v[i] * slope);
            dst[(size_t)j * p.L] = pack8(v);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[accb]);
      ++t_cnt;
    }
  } else if (warp >= EPI_WARP0 && !(dbg & 256)) {
    // ===== epilogue: quarter q = warp % 4 owns TMEM lanes 32q..32q+31 (one output row per lane);
    // the EPI_GROUPS warps of a quarter interleave over the (row tile m, 16-column block cb) items.
    const int quarter = warp & 3, grp = (warp - EPI_WARP0) >> 2;
    const int n_cb = p.NT / EPI_COLS, items = MT * n_cb, cb_shift = 31 - __clz(n_cb);
    const int CP = p.Cout_real / 8;
    const int cout_shift = 31 - __clz(p.Cout_real);       // Cout_real is a power of two (checked on the host)
    constexpr int NCH = EPI_COLS / 8;                     // 8-channel chunks per item
    const bool io = !(dbg & (2 | 2048));
    uint32_t t_cnt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(tile, p);
      const int t0 = tc.rt * (BM * MT);
      if (t0 >= row_end(p, tc.b)) continue;
      const int n0 = tc.ntile * p.NT;
      const uint32_t accb = t_cnt & 1u;
      const float* bias = p.bias ? p.bias + n0 : nullptr;
      const float* bbias = p.bbias ? p.bbias + (size_t)tc.b * p.bbias_ld + n0 : nullptr;
      const size_t plane0 = (size_t)tc.b * CP;
      const int qbase = t0 + quarter * 32 + lane;
      if (!freerun) mbar_wait(&acc_full[accb], (t_cnt >> 1) & 1u);
      tcgen05_fence_after();
      if (tid == EPI_WARP0 * 32) trace(dbg, 3, 1, (int)t_cnt);
#pragma unroll 1
      for (int it = (dbg & 32) ? items : grp; it < items; it += EPI_GROUPS) {
        const int m = it >> cb_shift, cb = it & (n_cb - 1);
        const int q = qbase + m * BM;
        const bool ok = q < p.L && io;
        const int c0 = cb * EPI_COLS;                      // column inside the tile
        const int n = n0 + c0;                             // GEMM column -> (phase r, channel co)
        const int r = n >> cout_shift, co = n & (p.Cout_real - 1);
        const size_t off0 = (plane0 + (co >> 3)) * p.L_out + (size_t)(q < p.L ? q : 0) * p.row_mul + r;
        // independent loads first: accumulate-input, bias
        uint4 ain[2 * NCH];
        if (ok && p.accin16) {
#pragma unroll
          for (int j = 0; j < NCH; ++j) ain[j] = *(reinterpret_cast<const uint4*>(p.accin16) + off0 + (size_t)j * p.L_out);
        }
        if (ok && p.accin32) {
#pragma unroll
          for (int j = 0; j < NCH; ++j) {
            ain[2 * j] = *(reinterpret_cast<const uint4*>(p.accin32) + (off0 + (size_t)j * p.L_out) * 2);
            ain[2 * j + 1] = *(reinterpret_cast<const uint4*>(p.accin32) + (off0 + (size_t)j * p.L_out) * 2 + 1);
          }
        }
        float bv[EPI_COLS];
#pragma unroll
        for (int i = 0; i < EPI_COLS; ++i) bv[i] = 0.f;
        if (p.bias_smem) {
#pragma unroll
          for (int i = 0; i < EPI_COLS / 4; ++i) {
            const float4 b4 = *(reinterpret_cast<const float4*>(bias_s + n0 + c0) + i);
            bv[4 * i] = b4.x; bv[4 * i + 1] = b4.y; bv[4 * i + 2] = b4.z; bv[4 * i + 3] = b4.w;
          }
        } else if (bias) {
#pragma unroll
          for (int i = 0; i < EPI_COLS / 4; ++i) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c0) + i);
            bv[4 * i] = b4.x; bv[4 * i + 1] = b4.y; bv[4 * i + 2] = b4.z; bv[4 * i + 3] = b4.w;
          }
        }
        if (bbias) {
#pragma unroll
          for (int i = 0; i < EPI_COLS / 4; ++i) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bbias + c0) + i);
            bv[4 * i] += b4.x; bv[4 * i + 1] += b4.y; bv[4 * i + 2] += b4.z; bv[4 * i + 3] += b4.w;
          }
        }
        // residual item from the shared ring (the loader runs a ring ahead of the MMAs)
        uint4 rcur[2 * NCH];
        if (p.r_slots > 0 && !(dbg & 16384)) {
          const int n_rb = p.NT / p.res_cols;
          const uint32_t sidx = (t_cnt * (uint32_t)MT + (uint32_t)m) * (uint32_t)n_rb + (uint32_t)(c0 / p.res_cols);
          const uint32_t slot = sidx % (uint32_t)p.r_slots;
          if (!freerun) mbar_wait(&r_full[slot], (sidx / (uint32_t)p.r_slots) & 1u);
          const uint32_t esz = p.res32 ? 32u : 16u;
          const uint8_t* rs = r_ring + (size_t)slot * p.r_slot_bytes + (size_t)((c0 % p.res_cols) >> 3) * (BM * esz);
          const int rrow = quarter * 32 + lane;
          if (p.res32) {
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
              rcur[2 * j] = *reinterpret_cast<const uint4*>(rs + j * (BM * 32) + rrow * 32);
              rcur[2 * j + 1] = *reinterpret_cast<const uint4*>(rs + j * (BM * 32) + rrow * 32 + 16);
            }
          } else {
#pragma unroll
            for (int j = 0; j < NCH; ++j) rcur[j] = *reinterpret_cast<const uint4*>(rs + j * (BM * 16) + rrow * 16);
          }
          // release the slot with an arrive that is data-dependent on the loaded registers: the refill is an
          // async-proxy write and a plain arrive does not wait for this warp's outstanding shared loads
          uint32_t dep = rcur[0].x ^ rcur[0].w ^ rcur[1].x ^ rcur[1].w;
          if (p.res32) dep ^= rcur[2].x ^ rcur[2].w ^ rcur[3].x ^ rcur[3].w;
          asm volatile("" : "+r"(dep));
          __syncwarp();
          if (lane == 0 && !freerun) mbar_arrive_dep(&r_empty[slot], dep);
        }
        uint32_t acc[EPI_COLS];
#pragma unroll
        for (int i = 0; i < EPI_COLS; ++i) acc[i] = 0u;
        if (!(dbg & 32768))
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + accb * (uint32_t)(MT * p.NT) +
                      (uint32_t)(m * p.NT + c0), acc);
        float sv[NZ_MAXK];       // source samples under the noise conv's taps of this lane's output row
        if (p.nz_src && ok) {
          const float* sp = p.nz_src + (size_t)tc.b * p.nz_len;
          const int s0 = (q * p.row_mul + r) * p.nz_stride - p.nz_pad;
#pragma unroll
          for (int jj = 0; jj < NZ_MAXK; ++jj) {
            const int sn = s0 + jj;
            sv[jj] = (jj < p.nz_k && sn >= 0 && sn < p.nz_len) ? __ldg(sp + sn) : 0.f;
          }
        }
        if (!(dbg & 65536))
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const size_t off = off0 + (size_t)j * p.L_out;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[j * 8 + i]) + bv[j * 8 + i];
          if (p.res16) {
            float rr[8];
            unpack8(rcur[j], rr);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += fminf(rr[i], rr[i] * p.res_inv);   // res_inv >= 1: undo lrelu
          }
          if (p.res32) {
            const uint4 q0 = rcur[2 * j], q1 = rcur[2 * j + 1];
            v[0] += __uint_as_float(q0.x); v[1] += __uint_as_float(q0.y);
            v[2] += __uint_as_float(q0.z); v[3] += __uint_as_float(q0.w);
            v[4] += __uint_as_float(q1.x); v[5] += __uint_as_float(q1.y);
            v[6] += __uint_as_float(q1.z); v[7] += __uint_as_float(q1.w);
          }
          if (p.out_scale != 1.f) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= p.out_scale;
          }
          if (p.nz_src && ok) {     // x + noise_convs[i](har_source): source taps in sv[], weights in shared memory
            const float* wb = nz_s + co + j * 8;
            {
              const float4 b0 = *reinterpret_cast<const float4*>(wb), b1 = *reinterpret_cast<const float4*>(wb + 4);
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
#pragma unroll
            for (int jj = 0; jj < NZ_MAXK; ++jj) {
              if (jj < p.nz_k) {
                const float4 w0 = *reinterpret_cast<const float4*>(wb + (size_t)(jj + 1) * p.Cout_real);
                const float4 w1 = *reinterpret_cast<const float4*>(wb + (size_t)(jj + 1) * p.Cout_real + 4);
                const float x = sv[jj];
                v[0] = fmaf(w0.x, x, v[0]); v[1] = fmaf(w0.y, x, v[1]);
                v[2] = fmaf(w0.z, x, v[2]); v[3] = fmaf(w0.w, x, v[3]);
                v[4] = fmaf(w1.x, x, v[4]); v[5] = fmaf(w1.y, x, v[5]);
                v[6] = fmaf(w1.z, x, v[6]); v[7] = fmaf(w1.w, x, v[7]);
              }
            }
          }
          if (ok) {
            if (p.accin16) {
              float o[8];
              unpack8(ain[j], o);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += o[i];
            }
            if (p.accin32) {
              const uint4 q0 = ain[2 * j], q1 = ain[2 * j + 1];
              v[0] += __uint_as_float(q0.x); v[1] += __uint_as_float(q0.y);
              v[2] += __uint_as_float(q0.z); v[3] += __uint_as_float(q0.w);
              v[4] += __uint_as_float(q1.x); v[5] += __uint_as_float(q1.y);
              v[6] += __uint_as_float(q1.z); v[7] += __uint_as_float(q1.w);
            }
            if (p.out32) {
              float4* o = reinterpret_cast<float4*>(p.out32) + off * 2;
              o[0] = make_float4(v[0], v[1], v[2], v[3]);
              o[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
            if (p.out16) {
              if (p.out16_slope != 1.f) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], v[i] * p.out16_slope);   // slope <= 1
              }
              const uint4 hi = pack8(v);
              *(reinterpret_cast<uint4*>(p.out16) + off) = hi;
              if (p.out_lo) {       // hi/lo stream of the last stage: the f16 rounding remainder
                float hv[8];
                unpack8(hi, hv);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] -= hv[i];
                *(reinterpret_cast<uint4*>(p.out_lo) + off) = pack8(v);
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[accb]);
      if (tid == EPI_WARP0 * 32) trace(dbg, 3, 2, (int)t_cnt);
      ++t_cnt;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) trace(dbg, 4, 1, 0);
  if (warp == MMA_WARP) tcgen05_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

struct Plan {
  int MT, NT, KC, PG, n_groups, a_slots, a_slot_bytes, stages, stage_bytes, plane_bytes, tmem_cols, resident;
  int r_slots, r_slot_bytes, res_cols, bias_smem;
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
  const int bias_smem = a.bias && a.N <= 4096 ? 1 : 0;
  const size_t fixed = 1024 + 1024 + (bias_smem ? 4 * (size_t)a.N + 16 : 0) +      // alignment slack, barriers, bias
                       (a.nz_src ? 4 * (size_t)(a.nz_k + 1) * a.Cout_real : 0);     // fused noise conv: bias + [k][C] weights
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
      const bool has_res = a.res16 || a.res32;
      const int res_cols = NT < 64 ? NT : 64;
      const int r_slot_bytes = has_res ? BM * (res_cols / 8) * (a.res32 ? 32 : 16) : 0;
      const int items = MT * (NT / res_cols);   // residual slots per tile
      int r_slots = has_res ? 2 : 0;
      if (fixed + 2 * (size_t)a_slot_bytes + w_min + (size_t)r_slots * r_slot_bytes > budget) continue;
      size_t left = budget - fixed - 2 * (size_t)a_slot_bytes - w_min - (size_t)r_slots * r_slot_bytes;
      int a_slots = 2, stages = resident ? n_chunks * a.K : 2;
      const int total_w = n_chunks * a.K;
      auto grow = [&](int* v, int cap, size_t unit) {
        while (*v < cap && left >= unit) {
          ++*v;
          left -= unit;
        }
      };
      const int max_stages = forced_stages > 0 ? forced_stages : 6;
      const int max_slots = forced_slots > 0 ? forced_slots : 4;
      if (!resident) grow(&stages, std::min(std::min(4, max_stages), total_w), stage_bytes);
      if (has_res) grow(&r_slots, items + 1, r_slot_bytes);      // one tile of residual slots ahead
      grow(&a_slots, std::min(3, max_slots), a_slot_bytes);
      if (!resident) grow(&stages, std::min(max_stages, total_w), stage_bytes);
      grow(&a_slots, max_slots, a_slot_bytes);
      if (has_res) grow(&r_slots, 2 * items, r_slot_bytes);
      int tm = 32;
      while (tm < 2 * MT * NT) tm <<= 1;
      *out = Plan{MT, NT, KC, PG, n_groups, a_slots, a_slot_bytes, stages, stage_bytes, plane_bytes, tm,
                  resident ? 1 : 0, r_slots, r_slot_bytes, res_cols, bias_smem,
                  fixed + (size_t)a_slots * a_slot_bytes + (size_t)stages * stage_bytes +
                      (size_t)r_slots * r_slot_bytes};
      return true;
    }
  }
  return false;
}

template <int MT, int KC16, int EPI, bool DBG, bool SWAP = false>
cudaError_t launch_t(const PlaneConvArgs& a, const Plan& pl, cudaStream_t s) {
  CUtensorMap wmap;
  if (!get_weight_map(a.w16, a.Cin, a.N, a.K, pl.KC, pl.NT, &wmap)) return cudaErrorNotSupported;
  PlaneParams p;
  p.x = a.x; p.L = a.L; p.lens = a.lens; p.in_mask = a.in_mask;
  p.tlen = a.tlen; p.len_mul = a.len_mul;
  p.B = a.B; p.Cin = a.Cin; p.K = a.K; p.dil = a.dil; p.pad = a.pad;
  p.bias = a.bias; p.bbias = a.bbias; p.bbias_ld = a.bbias_ld; p.tapmask = a.tapmask;
  p.N = a.N; p.NT = pl.NT; p.n_ntiles = a.N / pl.NT;
  p.n_row_tiles = (a.L + BM * MT - 1) / (BM * MT);
  p.total_tiles = p.n_row_tiles * a.B * p.n_ntiles;
  p.Cout_real = a.Cout_real; p.row_mul = a.row_mul; p.L_out = a.L * a.row_mul;
  p.res16 = a.res16; p.res_inv = a.res_inv; p.res32 = a.res32;
  p.accin16 = a.accin16; p.accin32 = a.accin32;
  p.out16 = a.out16; p.out16_slope = a.out16_slope; p.out32 = a.out32; p.out_scale = a.out_scale;
  p.nz_src = a.nz_src; p.nz_w = a.nz_w; p.nz_b = a.nz_b; p.nz_k = a.nz_k; p.nz_stride = a.nz_stride;
  p.nz_pad = a.nz_pad; p.nz_len = a.nz_len; p.out_lo = a.out_lo;
  p.plane_bytes = pl.plane_bytes; p.KC = pl.KC; p.PG = pl.PG; p.n_groups = pl.n_groups;
  p.a_slots = pl.a_slots; p.a_slot_bytes = pl.a_slot_bytes; p.stages = pl.stages;
  p.stage_bytes = pl.stage_bytes; p.resident = pl.resident; p.tmem_cols = pl.tmem_cols;
  p.r_slots = pl.r_slots; p.r_slot_bytes = pl.r_slot_bytes; p.res_cols = pl.res_cols; p.bias_smem = pl.bias_smem;
  p.idesc = SWAP ? make_idesc(pl.NT, BM * MT) : make_idesc(BM, pl.NT);
  static const int dbg = env_int("PG_PLANES_DEBUG", 0);
  p.debug = dbg;
  static const int nz_prefetch = env_int("PG_NZ_PREFETCH", 1);
  p.nz_prefetch = nz_prefetch;
  static DeviceOnce once;
  if (cudaError_t e = ensure_dyn_smem(conv_planes_kernel<MT, KC16, EPI, DBG, SWAP>, once, 227 * 1024)) return e;
  int grid = device_sm_count();
  if (a.grid_cap > 0 && grid > a.grid_cap) grid = a.grid_cap;
  if (grid > p.total_tiles) grid = p.total_tiles;
  if (p.debug & 8) {
    unsigned int zero = 0;
    cudaMemcpyToSymbol(g_ptrace_n, &zero, sizeof(zero));
  }
  if (cudaError_t e = launch_pdl(conv_planes_kernel<MT, KC16, EPI, DBG, SWAP>, dim3(grid), dim3(NTHREADS), pl.smem, s, wmap, p)) return e;
  if (p.debug & 8) {
    cudaDeviceSynchronize();
    static unsigned long long host[4096];
    unsigned int n = 0;
    cudaMemcpyFromSymbol(&n, g_ptrace_n, sizeof(n));
    cudaMemcpyFromSymbol(host, g_ptrace, sizeof(host));
    if (n > 4096) n = 4096;
    unsigned long long t0 = ~0ull;
    for (unsigned i = 0; i < n; ++i) t0 = host[i] >> 20 < t0 ? host[i] >> 20 : t0;
    fprintf(stderr, "TRACE MT=%d NT=%d resident=%d a_slots=%d stages=%d r_slots=%d grid=%d tiles=%d smem=%zu\n", MT,
            pl.NT, pl.resident, pl.a_slots, pl.stages, pl.r_slots, grid, p.total_tiles, pl.smem);
    for (unsigned i = 0; i < n; ++i)
      fprintf(stderr, "TR %llu role=%llu ev=%llu tile=%llu\n", (host[i] >> 20) - t0, (host[i] >> 16) & 15,
              (host[i] >> 12) & 15, host[i] & 0xFFF);
  }
  return cudaGetLastError();
}

}  // namespace

int plane_pick_nt(int n) { return pick_nt(n); }

bool plane_conv_supported(const PlaneConvArgs& a) {
  if (!a.x || !a.w16 || (!a.out16 && !a.out32)) return false;
  if (a.Cin % 32 || a.Cin < 32 || a.N % 32) return false;
  if (a.K < 1 || a.K > 32 || a.L <= 0 || a.B <= 0) return false;
  if (a.row_mul < 1 || a.Cout_real % EPI_COLS || (a.Cout_real & (a.Cout_real - 1)) || a.Cout_real * a.row_mul != a.N)
    return false;
  if (a.in_mask && !a.lens) return false;
  if ((a.res16 || a.res32) && a.row_mul != 1) return false;
  if (a.nz_src && (!a.nz_w || !a.nz_b || a.nz_k < 1 || a.nz_k > NZ_MAXK || a.nz_stride < 1 || a.Cout_real % 8)) return false;
  if (a.out_lo && !a.out16) return false;
  if (a.res16 && a.res32) return false;
  Plan pl;
  return make_plan(a, &pl);
}

int plane_conv_mt(const PlaneConvArgs& a) {
  Plan pl;
  return make_plan(a, &pl) ? pl.MT : 0;
}

cudaError_t launch_conv_planes(const PlaneConvArgs& a, cudaStream_t s) {
  Plan pl;
  if (!plane_conv_supported(a) || !make_plan(a, &pl)) return cudaErrorInvalidValue;
  static const bool dbg = env_int("PG_PLANES_DEBUG", 0) != 0;
  // straight-line epilogues for the plain ResBlock shapes
  int epi = EPI_GENERIC;
  if (!dbg && pl.bias_smem && a.N == pl.NT && a.row_mul == 1 && !a.bbias && !a.res32 && !a.accin32 && !a.nz_src &&
      !a.out_lo && !a.out32 && a.out16 && a.out16_slope <= 1.f && a.res_inv >= 1.f) {
    if (!a.accin16 && a.out_scale == 1.f) epi = a.res16 ? EPI_C2 : EPI_C1;
    else if (a.res16) epi = EPI_C3;
  }
  // polyphase upsamplers / bias-only convs over several Cout tiles (also the tensor-core noise conv: + accin16)
  if (epi == EPI_GENERIC && !dbg && pl.bias_smem && !a.bbias && !a.res16 && !a.res32 && !a.accin32 && !a.out32 &&
      a.out16 && a.out16_slope <= 1.f && a.out_scale == 1.f)
    epi = EPI_UPS;     // optional: fused noise conv (nz_*), accin16, hi/lo output
  // operand-swapped variant for the C = 128 ResBlock convs: +13 % on the conv1 shapes (1190 -> 1345 TFLOP/s at
  // k = 11, profiles/r02b), neutral on conv2 -- on by default, PG_FLAG_NO_PLANES_SWAP / PG_PLANES_SWAP=0 is the twin
  static const bool swap_on = env_int("PG_PLANES_SWAP", 1) != 0;
  if (swap_on && a.swap && (epi == EPI_C1 || epi == EPI_C2) && pl.MT == 2 && pl.NT == 128 && pl.KC == 64 && pl.res_cols == 64 &&
      a.Cout_real == 128)
    return epi == EPI_C1 ? launch_t<2, 4, EPI_C1, false, true>(a, pl, s) : launch_t<2, 4, EPI_C2, false, true>(a, pl, s);
#define PG_DISPATCH(MT_, KC16_)                                                        \
  return dbg ? launch_t<MT_, KC16_, EPI_GENERIC, true>(a, pl, s)                       \
             : (epi == EPI_C1 ? launch_t<MT_, KC16_, EPI_C1, false>(a, pl, s)          \
                              : (epi == EPI_C2 ? launch_t<MT_, KC16_, EPI_C2, false>(a, pl, s) \
                              : (epi == EPI_C3 ? launch_t<MT_, KC16_, EPI_C3, false>(a, pl, s) \
                              : (epi == EPI_UPS ? launch_t<MT_, KC16_, EPI_UPS, false>(a, pl, s) \
                                               : launch_t<MT_, KC16_, EPI_GENERIC, false>(a, pl, s)))))
  if (pl.KC == 64) {
    switch (pl.MT) {
      case 1: PG_DISPATCH(1, 4);
      case 2: PG_DISPATCH(2, 4);
      case 4: PG_DISPATCH(4, 4);
    }
  } else {
    switch (pl.MT) {
      case 1: PG_DISPATCH(1, 2);
      case 2: PG_DISPATCH(2, 2);
      case 4: PG_DISPATCH(4, 2);
    }
  }
#undef PG_DISPATCH
  return cudaErrorInvalidValue;
}

}  // namespace pg
