// Fused ResBlock pair on channel planes (sm_100a): one dilated conv, leaky-ReLU, one plain conv and
// the residual add of reference rvc/lib/algorithm/residuals.py:47-52
//     xt = c1(lrelu(x)) ; xt = c2(lrelu(xt)) ; x = xt + x
// in ONE kernel for the narrow decoder stages (C = 32 / 64), where the unfused layers are HBM-bound
// (arithmetic intensity 32-350 flop/B against a ridge of 218).  The intermediate never leaves the SM
// and the residual is read from the input window that is already in shared memory:
// HBM traffic per pair drops from 5 tensor passes to 2.
//
// Per tile of MT*128 intermediate rows (MO = MT*128 - (K-1) output rows, halo recomputed):
//   loader warp      bulk-copies the input window (MT*128 + (K-1)*dil rows of every 8-channel plane)
//   MMA lane         GEMM 1: acc1[rows][C] = sum_tap W1[tap] * A[row + tap*dil]      (tcgen05, TMEM)
//   8 warps  (E1)    acc1 -> +bias1 -> leaky-ReLU -> zero outside [0, L) -> f16 planes in SHARED memory
//   MMA lane         GEMM 2: acc2[rows][C] = sum_tap W2[tap] * tmp[row + tap]
//   8 warps  (E2)    acc2 -> +bias2 + residual (window rows in shared memory, or the fp32 stream through a
//                    bulk-copy ring) -> scale / accumulate -> L-form f16 and/or fp32 planes in HBM
// Both weight sets stay resident in shared memory (TMA, 128B/64B swizzle); accumulators and the
// intermediate are double-buffered so GEMM 1 of tile i+1 overlaps E1 of tile i and GEMM 2 / E2 of i-1.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include <algorithm>

#include "pg_common.cuh"
#include "pg_umma.cuh"

namespace pg {

namespace {

using namespace umma;

constexpr int BM = 128;
constexpr int LOAD_WARP = 0, TMA_WARP = 1, MMA_WARP = 2, RES_WARP = 3, E1_WARP0 = 4, E1_WARPS = 8, E2_WARP0 = 12,
              E2_WARPS = 8;
constexpr int NTHREADS = (E2_WARP0 + E2_WARPS) * 32;
constexpr int ECOLS = 16;   // accumulator columns per epilogue item

struct PairParams {
  const float* post_w; float* post_part; float post_slope;   // conv_post partials (PairConvArgs)
  const __half* x; int L, B, K, dil;
  const int* tlen; int len_mul;        // hard end of row b: tlen[b]*len_mul (nullptr = L)
  const float* bias1; const float* bias2;
  const float* res32; float res_inv;
  const __half* x_lo; __half* out_lo;     // HL stream: remainder planes of the input / output
  const __half* accin16; const float* accin32;
  __half* out16; float out16_slope; float* out32; float out_scale;
  int MO, n_row_tiles, total_tiles, h1, h2;
  int plane_bytes, a_slots, a_slot_bytes;
  int tmp_plane_bytes, tmp_bytes;
  int w_tile_bytes;
  int r_slots, r_slot_bytes;
  int tmem_cols;
  int lag;      // GEMM 2 trails GEMM 1 by this many tiles (1 .. depth-1)
  int depth;    // D: tiles in flight between GEMM 1 and E2 (accumulator pairs and intermediates are D-buffered)
  uint32_t idesc;
  int no_ring;  // PG_PAIR_NORING (debug): fp32 residual read straight from global memory in E2
  int no_pre;   // PG_PAIR_NOPRE (debug): accumulate input loaded at use instead of prefetched
};

// hard end of batch row b; tiles whose first output row lies at or beyond it are skipped by every role
__device__ __forceinline__ int row_end(const PairParams& p, int b) {
  return p.tlen ? min(p.L, p.tlen[b] * p.len_mul) : p.L;
}

__device__ unsigned long long g_pair_dbg[16];   // diagnostics (PG_PAIR_NOPRE=3): early vs late accumulate-input reads

// accumulate-input load: mode 2 (PG_PAIR_NOPRE=2, diagnostic) bypasses L1 (ld.global.cg)
__device__ __forceinline__ uint4 ld_acc(const uint4* p, int mode) {
  return mode == 2 ? __ldcg(p) : *p;
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

// RES_X: the residual is the raw value of the f16 input stream (else: the fp32 stream through the ring)
// PLAIN (RES_X only): the 2 of 3 pairs of a ResBlock that just write y16 = lrelu(conv2 + bias2 + raw(x)) -- no
// accumulate input, unit scale, no fp32 output -- get a straight-line E2 (the generic one spends ~360
// instructions per 32x16 item on run-time options and re-derived addresses, which made E2 the issue-bound
// stage of the C = 64 pairs: profiles/r02d_pair64k3)
// HL (with RES_X): the stream is carried as TWO f16 planes, x = hi + lo (hi is the conv operand, lo its rounding
// remainder: 2^-22 relative, fp32-class), so the last decoder stage keeps its residual stream at full precision
// without a separate fp32 copy -- 256 instead of 384 bytes per sample and pair.
// EPIM selects the E2 code: 0 generic (every run-time option), 1 plain pair, 2 last pair of a ResBlock (scaled by
// 1/n_kernels; fp32 mean planes out for HL, f16 out otherwise), 3 the same plus the running mean as accumulate input
template <int MT, int C, bool RES_X, int EPIM, bool HL>
__global__ void __launch_bounds__(NTHREADS, 1)
pair_planes_kernel(const __grid_constant__ CUtensorMap wmap1, const __grid_constant__ CUtensorMap wmap2,
                   const PairParams p) {
  pdl_launch_dependents();     // the next kernel of the stream may launch and run its prologue (pg_common.cuh)
  constexpr int PLANES = C / 8;
  constexpr int KC16 = C / 16;                 // one K chunk = all C input channels
  constexpr int NCB = C / ECOLS, ITEMS = MT * NCB;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w1 = smem;                                              // [K] tiles, GEMM 1
  uint8_t* w2 = w1 + (size_t)p.K * p.w_tile_bytes;                 // [K] tiles, GEMM 2
  uint8_t* a_ring = w2 + (size_t)p.K * p.w_tile_bytes;
  uint8_t* tmp = a_ring + (size_t)p.a_slots * p.a_slot_bytes;      // [D] intermediate planes
  const uint32_t D = (uint32_t)p.depth;
  uint8_t* r_ring = tmp + (size_t)D * p.tmp_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(r_ring + (size_t)p.r_slots * p.r_slot_bytes);
  uint64_t* a_full = bars;                    // [a_slots]
  uint64_t* a_empty = a_full + p.a_slots;     // [a_slots]  GEMM 1 commit
  uint64_t* acc1_full = a_empty + p.a_slots;  // [D]
  uint64_t* acc1_empty = acc1_full + D;       // [D]  E1 warps
  uint64_t* tmp_full = acc1_empty + D;        // [D]  E1 warps
  uint64_t* tmp_empty = tmp_full + D;         // [D]  GEMM 2 commit
  uint64_t* acc2_full = tmp_empty + D;        // [D]
  uint64_t* acc2_empty = acc2_full + D;       // [D]  E2 warps
  uint64_t* w_ready = acc2_empty + D;
  uint64_t* r_full = w_ready + 1;             // [r_slots]
  uint64_t* r_empty = r_full + p.r_slots;     // [r_slots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_empty + p.r_slots);
  float* bias_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));   // [2][C]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < p.a_slots; ++s) {
      mbar_init(&a_full[s], 1);
      // The window slot is released by GEMM 1's commit alone.  E2 used to read the residual rows from the
      // window and so held the slot through E1 / GEMM 2 / E2: with 3 slots that left ~1/3 of the ring's
      // bytes in flight towards HBM and the kernel latency-bound (ncu r02d: MMA lane and loader both spinning
      // on the ring, DRAM 22 %, tensor pipe 19 % at C = 64 k = 3).  E2 now re-reads its residual rows from
      // global memory -- they were fetched microseconds ago, so they come from L2.
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < (int)D; ++s) {
      mbar_init(&acc1_full[s], 1);
      mbar_init(&acc1_empty[s], E1_WARPS);
      mbar_init(&tmp_full[s], E1_WARPS);
      mbar_init(&tmp_empty[s], 1);
      mbar_init(&acc2_full[s], 1);
      mbar_init(&acc2_empty[s], E2_WARPS);
    }
    mbar_init(w_ready, 1);
    for (int s = 0; s < p.r_slots; ++s) {
      mbar_init(&r_full[s], 1);
      mbar_init(&r_empty[s], 4 * NCB);      // every (quarter warp, item) that reads the 128-row slot
    }
    fence_barrier_init();
  }
  if (tid < 2 * C) bias_s[tid] = tid < C ? p.bias1[tid] : p.bias2[tid - C];
  float* post_s = bias_s + 2 * C;      // [K_POST][C] conv_post weights (fused post step)
  if (p.post_part)
    for (int i = tid; i < K_POST * C; i += NTHREADS) post_s[i] = p.post_w[i];
  if (warp == MMA_WARP) tcgen05_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  if (warp == TMA_WARP && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap2) : "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // activations in global memory belong to the predecessor until it has completed; the weight producer and the MMA
  // issuer touch constants / shared memory / TMEM only and run ahead
  if (warp != TMA_WARP && warp != MMA_WARP) pdl_wait();
  const int rows = BM * MT + (p.K - 1) * p.dil;        // input window rows
  const uint32_t acc1_col = 0u, acc2_col = D * MT * C;  // TMEM columns of the two accumulator sets

  if (warp == LOAD_WARP) {
    // ===== input window loader (bulk copies + zero fill outside [0, L)) =====
    uint32_t cnt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int rt = tile % p.n_row_tiles, b = tile / p.n_row_tiles;
      const int t_end = row_end(p, b);
      if (rt * p.MO >= t_end) continue;
      const int tstart = rt * p.MO - p.h2 - p.h1;
      const int lo = min(max(-tstart, 0), rows);
      const int hi = min(max(t_end - tstart, lo), rows);
      const int nz = lo + (rows - hi);
      const __half* xb = p.x + (size_t)b * PLANES * p.L * 8;
      const uint32_t slot = cnt % (uint32_t)p.a_slots;
      mbar_wait(&a_empty[slot], ((cnt / (uint32_t)p.a_slots) & 1u) ^ 1u);
      uint8_t* dst = a_ring + (size_t)slot * p.a_slot_bytes;
      if (nz > 0) {
        for (int i = lane; i < PLANES * nz; i += 32) {
          const int pl = i / nz, z = i - pl * nz;
          const int r = z < lo ? z : hi + (z - lo);
          *reinterpret_cast<uint4*>(dst + (size_t)pl * p.plane_bytes + (size_t)r * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
      }
      __syncwarp();
      const uint32_t bytes = (uint32_t)(hi - lo) * 16u;
      if (lane == 0) mbar_expect_tx(&a_full[slot], bytes * (uint32_t)PLANES);
      __syncwarp();
      if (bytes && lane < PLANES)
        bulk_g2s(dst + (size_t)lane * p.plane_bytes + (size_t)lo * 16,
                 xb + ((size_t)lane * p.L + (tstart + lo)) * 8, bytes, &a_full[slot]);
      ++cnt;
    }
  } else if (warp == TMA_WARP) {
    // ===== both weight sets, once =====
    if (elect_one()) {
      mbar_expect_tx(w_ready, (uint32_t)(2 * p.K * p.w_tile_bytes));
      for (int tap = 0; tap < p.K; ++tap) {
        tma_load_2d(w1 + (size_t)tap * p.w_tile_bytes, &wmap1, w_ready, 0, tap * C);
        tma_load_2d(w2 + (size_t)tap * p.w_tile_bytes, &wmap2, w_ready, 0, tap * C);
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    // ===== MMA issuer (one elected lane): GEMM 1 of tile i, then GEMM 2 of tile i-1 =====
    if (elect_one()) {
      constexpr uint32_t w_layout = C == 64 ? LAYOUT_SW128 : LAYOUT_SW64;
      constexpr uint32_t w_sbo = 8u * C * 2u;
      const uint64_t adesc0 = make_desc(0u, (uint32_t)p.plane_bytes, 128u, LAYOUT_NONE);
      const uint64_t tdesc0 = make_desc(0u, (uint32_t)p.tmp_plane_bytes, 128u, LAYOUT_NONE);
      const uint64_t bdesc0 = make_desc(0u, 0u, w_sbo, w_layout);
      const uint32_t a_hi = (uint32_t)(adesc0 >> 32), t_hi = (uint32_t)(tdesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
      const uint32_t a_lo0 = (uint32_t)adesc0, t_lo0 = (uint32_t)tdesc0, b_lo0 = (uint32_t)bdesc0;
      const uint32_t plane2 = 2u * ((uint32_t)p.plane_bytes >> 4), tplane2 = 2u * ((uint32_t)p.tmp_plane_bytes >> 4);
      const uint32_t w1_units = smem_u32(w1) >> 4, w2_units = smem_u32(w2) >> 4, wt_units = (uint32_t)p.w_tile_bytes >> 4;
      const uint32_t a_units0 = smem_u32(a_ring) >> 4, slot_units = (uint32_t)p.a_slot_bytes >> 4;
      const uint32_t tmp_units0 = smem_u32(tmp) >> 4, tmp_units = (uint32_t)p.tmp_bytes >> 4;
      const uint32_t idesc = p.idesc;
      mbar_wait(w_ready, 0);
      tcgen05_fence_after();
      int n_my = 0;     // live tiles of this CTA (dead ones are skipped by every role)
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x)
        n_my += (tile % p.n_row_tiles) * p.MO < row_end(p, tile / p.n_row_tiles) ? 1 : 0;
      // GEMM 2 trails GEMM 1 by LAG tiles: E1 of a tile has LAG GEMM-1 times to finish before the tensor pipe
      // needs its intermediate (each mbarrier hand-off costs a few hundred cycles, more than a k = 3 GEMM)
      const int LAG = p.lag;
      for (int it = 0; it < n_my + LAG; ++it) {
        if (it < n_my) {   // GEMM 1 (dilated conv) of tile it
          const uint32_t slot = (uint32_t)it % (uint32_t)p.a_slots, b = (uint32_t)it % D;
          mbar_wait(&a_full[slot], ((uint32_t)it / (uint32_t)p.a_slots) & 1u);
          mbar_wait(&acc1_empty[b], (((uint32_t)it / D) & 1u) ^ 1u);
          tcgen05_fence_after();
          const uint32_t d_base = tmem_base + acc1_col + b * (uint32_t)(MT * C);
          const uint32_t a_base = a_lo0 + a_units0 + slot * slot_units;
#pragma unroll 1
          for (int tap = 0; tap < p.K; ++tap) {
            const uint32_t a_lo = a_base + (uint32_t)(tap * p.dil);
            const uint32_t w_lo = b_lo0 + w1_units + (uint32_t)tap * wt_units;
#pragma unroll
            for (int k16 = 0; k16 < KC16; ++k16)
#pragma unroll
              for (int m = 0; m < MT; ++m)
                umma_f16_lh(d_base + (uint32_t)(m * C), a_lo + (uint32_t)k16 * plane2 + (uint32_t)(m * BM), a_hi,
                            w_lo + 2u * (uint32_t)k16, b_hi, idesc, (uint32_t)(tap | k16));
          }
          tcgen05_commit(&acc1_full[b]);
          tcgen05_commit(&a_empty[slot]);
        }
        if (it >= LAG) {   // GEMM 2 (plain conv over the shared-memory intermediate) of tile it-LAG
          const uint32_t j = (uint32_t)(it - LAG), b = j % D;
          mbar_wait(&tmp_full[b], (j / D) & 1u);
          mbar_wait(&acc2_empty[b], ((j / D) & 1u) ^ 1u);
          tcgen05_fence_after();
          const uint32_t d_base = tmem_base + acc2_col + b * (uint32_t)(MT * C);
          const uint32_t t_base = t_lo0 + tmp_units0 + b * tmp_units;
#pragma unroll 1
          for (int tap = 0; tap < p.K; ++tap) {
            const uint32_t t_lo = t_base + (uint32_t)tap;
            const uint32_t w_lo = b_lo0 + w2_units + (uint32_t)tap * wt_units;
#pragma unroll
            for (int k16 = 0; k16 < KC16; ++k16)
#pragma unroll
              for (int m = 0; m < MT; ++m)
                umma_f16_lh(d_base + (uint32_t)(m * C), t_lo + (uint32_t)k16 * tplane2 + (uint32_t)(m * BM), t_hi,
                            w_lo + 2u * (uint32_t)k16, b_hi, idesc, (uint32_t)(tap | k16));
          }
          tcgen05_commit(&acc2_full[b]);
          tcgen05_commit(&tmp_empty[b]);
        }
      }
    }
  } else if (warp == RES_WARP && p.r_slots > 0 && !p.no_ring) {
    // ===== fp32 residual stream: one 128-row x C slot per row tile m =====
    const int pl = lane & 7, sub = lane >> 3;
    const int batch = min(min(MT, p.r_slots), 4);
    uint32_t s_cnt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int rt = tile % p.n_row_tiles, b = tile / p.n_row_tiles;
      const int o0 = rt * p.MO;
      if (o0 >= row_end(p, b)) continue;
      for (int base = 0; base < MT; base += batch) {
        const int m = base + sub;
        const bool mine = sub < batch && m < MT && pl < PLANES;
        const uint32_t sidx = s_cnt + (uint32_t)m;
        const uint32_t slot = sidx % (uint32_t)p.r_slots;
        const int row0 = o0 + m * BM;
        const int nrows = min(BM, p.L - row0);
        const uint32_t bytes = nrows > 0 ? (uint32_t)nrows * 32u : 0u;
        if (mine) mbar_wait(&r_empty[slot], ((sidx / (uint32_t)p.r_slots) & 1u) ^ 1u);
        __syncwarp();
        if (mine && pl == 0) mbar_expect_tx(&r_full[slot], (uint32_t)PLANES * bytes);
        __syncwarp();
        if (HL) {
          // hi/lo stream: the slot holds [hi planes][lo planes], 16 B per row each (same bytes as an fp32 slot);
          // lanes 0..PLANES-1 of the sub-group copy the hi planes, PLANES..2*PLANES-1 the lo planes
          const bool mine2 = sub < batch && m < MT && pl < 2 * PLANES;
          if (mine2 && bytes) {
            const int q = pl < PLANES ? pl : pl - PLANES;
            const __half* srcp = pl < PLANES ? p.x : p.x_lo;
            bulk_g2s(r_ring + (size_t)slot * p.r_slot_bytes + (size_t)pl * (BM * 16),
                     reinterpret_cast<const char*>(srcp) + (((size_t)b * PLANES + q) * p.L + row0) * 16, bytes / 2,
                     &r_full[slot]);
          }
        } else if (mine && bytes)
          bulk_g2s(r_ring + (size_t)slot * p.r_slot_bytes + (size_t)pl * (BM * 32),
                   reinterpret_cast<const char*>(p.res32) + (((size_t)b * PLANES + pl) * p.L + row0) * 32, bytes,
                   &r_full[slot]);
      }
      s_cnt += (uint32_t)MT;
    }
  } else if (warp >= E1_WARP0 && warp < E2_WARP0) {
    // ===== E1: acc1 -> lrelu(acc + bias1) -> f16 planes of the intermediate, in shared memory =====
    const int quarter = warp & 3, grp = (warp - E1_WARP0) >> 2;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc1_col;
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int rt = tile % p.n_row_tiles;
      const int t_end = row_end(p, tile / p.n_row_tiles);
      if (rt * p.MO >= t_end) continue;
      const int t_first = rt * p.MO - p.h2;            // time of intermediate row 0
      const uint32_t b = j % D;
      mbar_wait(&acc1_full[b], (j / D) & 1u);
      mbar_wait(&tmp_empty[b], ((j / D) & 1u) ^ 1u);
      tcgen05_fence_after();
      uint8_t* tb = tmp + (size_t)b * p.tmp_bytes;
#pragma unroll 1
      for (int it = grp; it < ITEMS; it += E1_WARPS / 4) {
        const int m = it / NCB, cb = it - m * NCB;
        const int r = m * BM + quarter * 32 + lane;
        const int t = t_first + r;
        const bool live = t >= 0 && t < t_end;
        uint32_t acc[ECOLS];
        tmem_ld16(lane_taddr + b * (uint32_t)(MT * C) + (uint32_t)(m * C + cb * ECOLS), acc);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float a = __uint_as_float(acc[jj * 8 + i]) + bias_s[cb * ECOLS + jj * 8 + i];
            v[i] = live ? fmaxf(a, a * 0.1f) : 0.f;
          }
          *reinterpret_cast<uint4*>(tb + (size_t)(cb * 2 + jj) * p.tmp_plane_bytes + (size_t)r * 16) = pack8(v);
        }
      }
      fence_proxy_async_smem();      // intermediate rows -> visible to the tensor core
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tmp_full[b]);
        mbar_arrive(&acc1_empty[b]);
      }
      ++j;
    }
  } else if (warp >= E2_WARP0 && EPIM > 0) {
    // ===== E2, straight-line forms: v = (acc2 + bias2 + raw(x)) [* scale + accin]; y16 = lrelu(v, slope) / y32 = v =====
    constexpr bool LASTP = EPIM >= 2, ACC = EPIM == 3, OUT32 = LASTP && HL;
    const float scale = p.out_scale;
    const int quarter = warp & 3, grp = (warp - E2_WARP0) >> 2;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc2_col;
    const float* bias2 = bias_s + C;
    const float rinv = p.res_inv, slope = p.out16_slope;
    const int L = p.L, MO = p.MO;
    constexpr int IPW = ITEMS / (E2_WARPS / 4);
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int rt = tile % p.n_row_tiles, bb = tile / p.n_row_tiles;
      const int o0 = rt * MO;
      if (o0 >= row_end(p, bb)) continue;
      const uint32_t b = j % D;
      const size_t plane_base = (size_t)bb * PLANES * L;
      const uint4* xin = reinterpret_cast<const uint4*>(p.x) + plane_base;
      uint4* yout = reinterpret_cast<uint4*>(p.out16) + plane_base;
      uint4* lout = reinterpret_cast<uint4*>(p.out_lo) + plane_base;
      // NOTE: the accumulate input is loaded per item AFTER the accumulator wait.  Fetching it for the whole tile
      // before that wait (as round 1 did to hide its latency behind GEMM 2) made ~1 % of the calls differ from
      // each other in one 32-row item (tools/diag_det.py; 0 of 500 with the loads here, in both stream forms,
      // with and without the wait hint, and the early and a late read compare equal when both are made: the
      // cause was not found, so the early fetch is gone).
      uint4 apre[ACC ? (HL ? 4 : 2) * IPW : 1];
      uint4 rpre[HL ? 1 : 2 * IPW];
      if (!HL) {
#pragma unroll
        for (int ii = 0; ii < IPW; ++ii) {     // residual rows: fetched (from L2) before waiting for the accumulator
          const int it = grp + ii * (E2_WARPS / 4);
          const int m = it / NCB, cb = it - m * NCB;
          const int o = m * BM + quarter * 32 + lane;
          const int t = o0 + o;
          if (o < MO && t < L) {
            const uint4* g = xin + (size_t)(cb * 2) * L + t;
            rpre[2 * ii] = g[0];
            rpre[2 * ii + 1] = g[L];
          }
        }
      }
      mbar_wait(&acc2_full[b], (j / D) & 1u);
      tcgen05_fence_after();
#pragma unroll
      for (int ii = 0; ii < IPW; ++ii) {
        const int it = grp + ii * (E2_WARPS / 4);
        const int m = it / NCB, cb = it - m * NCB;
        const int o = m * BM + quarter * 32 + lane;
        const int t = o0 + o;
        if (ACC && p.no_pre != 1 && o < MO && t < L) {      // running mean of the earlier ResBlocks, this item's rows
          const size_t off = plane_base + (size_t)(cb * 2) * L + t;
          if (HL) {
            const uint4* g = reinterpret_cast<const uint4*>(p.accin32) + off * 2;
            apre[4 * ii] = g[0];
            apre[4 * ii + 1] = g[1];
            apre[4 * ii + 2] = g[2 * (size_t)L];
            apre[4 * ii + 3] = g[2 * (size_t)L + 1];
          } else {
            const uint4* g = reinterpret_cast<const uint4*>(p.accin16) + off;
            apre[2 * ii] = g[0];
            apre[2 * ii + 1] = g[L];
          }
        }
        uint4 hq[2], lq[2];
        if (HL) {     // hi/lo rows of this item from the residual ring (the loader runs tiles ahead of the MMAs)
          const uint32_t sidx = j * (uint32_t)MT + (uint32_t)m;
          const uint32_t rslot = sidx % (uint32_t)p.r_slots;
          mbar_wait(&r_full[rslot], (sidx / (uint32_t)p.r_slots) & 1u);
          const uint8_t* rs = r_ring + (size_t)rslot * p.r_slot_bytes + (size_t)(cb * 2) * (BM * 16) +
                              (size_t)(quarter * 32 + lane) * 16;
          hq[0] = *reinterpret_cast<const uint4*>(rs);
          hq[1] = *reinterpret_cast<const uint4*>(rs + BM * 16);
          lq[0] = *reinterpret_cast<const uint4*>(rs + PLANES * (BM * 16));
          lq[1] = *reinterpret_cast<const uint4*>(rs + (PLANES + 1) * (BM * 16));
          uint32_t dep = hq[0].x ^ hq[1].x ^ lq[0].x ^ lq[1].x ^ hq[0].w ^ hq[1].w ^ lq[0].w ^ lq[1].w;
          asm volatile("" : "+r"(dep));
          __syncwarp();
          if (lane == 0) mbar_arrive_dep(&r_empty[rslot], dep);   // released once the loads have landed
        }
        uint32_t acc[ECOLS];
        tmem_ld16(lane_taddr + b * (uint32_t)(MT * C) + (uint32_t)(m * C + cb * ECOLS), acc);
        if (o < MO && t < L) {
          uint4* dst = yout + (size_t)(cb * 2) * L + t;
          float pd[K_POST];       // fused conv_post: this item's partial dot products
#pragma unroll
          for (int jp = 0; jp < K_POST; ++jp) pd[jp] = 0.f;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            float v[8], rr[8];
            unpack8(HL ? hq[jj] : rpre[HL ? 0 : 2 * ii + jj], rr);
            if (HL) {
              float ll[8];
              unpack8(lq[jj], ll);
#pragma unroll
              for (int i = 0; i < 8; ++i) rr[i] += ll[i];
            }
            const float4 b0 = *reinterpret_cast<const float4*>(bias2 + cb * ECOLS + jj * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(bias2 + cb * ECOLS + jj * 8 + 4);
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[jj * 8 + i]) + bv[i] + fminf(rr[i], rr[i] * rinv);
            if (LASTP) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] *= scale;
            }
            if (ACC && p.no_pre == 1) {      // (PG_PAIR_NOPRE=1) accumulate input loaded at the point of use
              const size_t off = plane_base + (size_t)(cb * 2 + jj) * L + t;
              if (HL) {
                apre[4 * ii + 2 * jj] = *(reinterpret_cast<const uint4*>(p.accin32) + off * 2);
                apre[4 * ii + 2 * jj + 1] = *(reinterpret_cast<const uint4*>(p.accin32) + off * 2 + 1);
              } else {
                apre[2 * ii + jj] = *(reinterpret_cast<const uint4*>(p.accin16) + off);
              }
            }
            if (ACC) {
              if (HL) {
                const uint4 q0 = apre[4 * ii + 2 * jj], q1 = apre[4 * ii + 2 * jj + 1];
                v[0] += __uint_as_float(q0.x); v[1] += __uint_as_float(q0.y);
                v[2] += __uint_as_float(q0.z); v[3] += __uint_as_float(q0.w);
                v[4] += __uint_as_float(q1.x); v[5] += __uint_as_float(q1.y);
                v[6] += __uint_as_float(q1.z); v[7] += __uint_as_float(q1.w);
              } else {
                float av[8];
                unpack8(apre[2 * ii + jj], av);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] += av[i];
              }
            }
            if (OUT32) {
              if (p.post_part) {     // conv_post on the finished mean, in fp32: 7 taps x this chunk's 8 channels
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], v[i] * p.post_slope);     // slope <= 1
                const float* wp = post_s + cb * ECOLS + jj * 8;
#pragma unroll
                for (int jp = 0; jp < K_POST; ++jp) {
                  const float4 w0 = *reinterpret_cast<const float4*>(wp + jp * C);
                  const float4 w1 = *reinterpret_cast<const float4*>(wp + jp * C + 4);
                  float a = pd[jp];
                  a = fmaf(w0.x, v[0], a); a = fmaf(w0.y, v[1], a); a = fmaf(w0.z, v[2], a); a = fmaf(w0.w, v[3], a);
                  a = fmaf(w1.x, v[4], a); a = fmaf(w1.y, v[5], a); a = fmaf(w1.z, v[6], a); a = fmaf(w1.w, v[7], a);
                  pd[jp] = a;
                }
                continue;
              }
              // the 1/3-mean that feeds conv_post stays fp32
              float4* o4 = reinterpret_cast<float4*>(p.out32) + (plane_base + (size_t)(cb * 2 + jj) * L + t) * 2;
              o4[0] = make_float4(v[0], v[1], v[2], v[3]);
              o4[1] = make_float4(v[4], v[5], v[6], v[7]);
              continue;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], v[i] * slope);
            const uint4 hi = pack8(v);
            dst[(size_t)jj * L] = hi;
            if (HL) {     // remainder of the f16 rounding, itself in f16
              float hv[8];
              unpack8(hi, hv);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] -= hv[i];
              (lout + (size_t)(cb * 2) * L + t)[(size_t)jj * L] = pack8(v);
            }
          }
          if (OUT32 && p.post_part) {      // one coalesced 4 B store per tap: part[half cb][tap][batch row][t]
#pragma unroll
            for (int jp = 0; jp < K_POST; ++jp)
              p.post_part[((size_t)(cb * K_POST + jp) * p.B + bb) * L + t] = pd[jp];
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc2_empty[b]);
      ++j;
    }
  } else if (warp >= E2_WARP0) {
    // ===== E2: acc2 -> + bias2 + residual -> scale / accumulate -> HBM =====
    const int quarter = warp & 3, grp = (warp - E2_WARP0) >> 2;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc2_col;
    const float* bias2 = bias_s + C;
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int rt = tile % p.n_row_tiles, bb = tile / p.n_row_tiles;
      const int o0 = rt * p.MO;
      if (o0 >= row_end(p, bb)) continue;
      const uint32_t b = j % D;
      const size_t plane_base = (size_t)bb * PLANES * p.L;
      // the accumulate input of this warp's items is fetched BEFORE waiting for the accumulator, so its
      // HBM latency hides behind GEMM 2 (registers: IPW items x 2 chunks, x2 for fp32)
      constexpr int IPW = ITEMS / (E2_WARPS / 4);
      constexpr bool PRE32 = IPW * 4 <= 8;
      uint4 pre[8];
      // early fetch of the accumulate input: OFF (see the note in the straight-line E2); PG_PAIR_NOPRE=4 re-enables it
      const bool pre16 = p.accin16 != nullptr && p.no_pre == 4, pre32 = PRE32 && p.accin32 != nullptr && p.no_pre == 4;
      if (pre16 || pre32) {
#pragma unroll
        for (int ii = 0; ii < IPW; ++ii) {
          const int it = grp + ii * (E2_WARPS / 4);
          const int m = it / NCB, cb = it - m * NCB;
          const int o = m * BM + quarter * 32 + lane;
          const int t = o0 + o;
          if (o < p.MO && t < p.L) {
            const size_t off = plane_base + (size_t)(cb * 2) * p.L + t;
            if (pre16) {
              pre[2 * ii] = ld_acc(reinterpret_cast<const uint4*>(p.accin16) + off, p.no_pre);
              pre[2 * ii + 1] = ld_acc(reinterpret_cast<const uint4*>(p.accin16) + off + p.L, p.no_pre);
            } else if (PRE32) {
              pre[(4 * ii) & 7] = ld_acc(reinterpret_cast<const uint4*>(p.accin32) + off * 2, p.no_pre);
              pre[(4 * ii + 1) & 7] = ld_acc(reinterpret_cast<const uint4*>(p.accin32) + off * 2 + 1, p.no_pre);
              pre[(4 * ii + 2) & 7] = ld_acc(reinterpret_cast<const uint4*>(p.accin32) + (off + p.L) * 2, p.no_pre);
              pre[(4 * ii + 3) & 7] = ld_acc(reinterpret_cast<const uint4*>(p.accin32) + (off + p.L) * 2 + 1, p.no_pre);
            }
          }
        }
      }
      // f16 stream: the residual (raw x of the output rows) is prefetched from global memory (L2) as well
      uint4 rpre[RES_X && !HL ? 2 * IPW : 1];
      if (RES_X && !HL) {
#pragma unroll
        for (int ii = 0; ii < IPW; ++ii) {
          const int it = grp + ii * (E2_WARPS / 4);
          const int m = it / NCB, cb = it - m * NCB;
          const int o = m * BM + quarter * 32 + lane;
          const int t = o0 + o;
          if (o < p.MO && t < p.L) {
            const uint4* g = reinterpret_cast<const uint4*>(p.x) + plane_base + (size_t)(cb * 2) * p.L + t;
            rpre[2 * ii] = g[0];
            rpre[2 * ii + 1] = g[p.L];
          }
        }
      }
      mbar_wait(&acc2_full[b], (j / D) & 1u);
      tcgen05_fence_after();
#pragma unroll
      for (int ii = 0; ii < IPW; ++ii) {
        const int it = grp + ii * (E2_WARPS / 4);
        const int m = it / NCB, cb = it - m * NCB;
        const int o = m * BM + quarter * 32 + lane;
        const int t = o0 + o;
        const bool ok = o < p.MO && t < p.L;
        uint4 rq[4];
        int ring_slot = -1;
        if (!RES_X && p.no_ring) {
          if (ok) {
            const uint4* g0 = reinterpret_cast<const uint4*>(p.res32) + (plane_base + (size_t)(cb * 2) * p.L + t) * 2;
            const uint4* g1 = reinterpret_cast<const uint4*>(p.res32) + (plane_base + (size_t)(cb * 2 + 1) * p.L + t) * 2;
            rq[0] = g0[0]; rq[1] = g0[1]; rq[2] = g1[0]; rq[3] = g1[1];
          }
        } else if (!RES_X) {
          const uint32_t sidx = j * (uint32_t)MT + (uint32_t)m;
          const uint32_t rslot = sidx % (uint32_t)p.r_slots;
          mbar_wait(&r_full[rslot], (sidx / (uint32_t)p.r_slots) & 1u);
          const uint8_t* rs = r_ring + (size_t)rslot * p.r_slot_bytes + (size_t)(cb * 2) * (BM * 32) +
                              (size_t)(quarter * 32 + lane) * 32;
          rq[0] = *reinterpret_cast<const uint4*>(rs);
          rq[1] = *reinterpret_cast<const uint4*>(rs + 16);
          rq[2] = *reinterpret_cast<const uint4*>(rs + BM * 32);
          rq[3] = *reinterpret_cast<const uint4*>(rs + BM * 32 + 16);
          ring_slot = (int)rslot;   // released below, once the loads have landed in registers
        } else if (HL) {   // ring slot = [hi planes][lo planes]: rq[0..1] = hi chunks, rq[2..3] = lo chunks
          const uint32_t sidx = j * (uint32_t)MT + (uint32_t)m;
          const uint32_t rslot = sidx % (uint32_t)p.r_slots;
          mbar_wait(&r_full[rslot], (sidx / (uint32_t)p.r_slots) & 1u);
          const uint8_t* rs = r_ring + (size_t)rslot * p.r_slot_bytes + (size_t)(cb * 2) * (BM * 16) +
                              (size_t)(quarter * 32 + lane) * 16;
          rq[0] = *reinterpret_cast<const uint4*>(rs);
          rq[1] = *reinterpret_cast<const uint4*>(rs + BM * 16);
          rq[2] = *reinterpret_cast<const uint4*>(rs + PLANES * (BM * 16));
          rq[3] = *reinterpret_cast<const uint4*>(rs + (PLANES + 1) * (BM * 16));
          ring_slot = (int)rslot;
        } else {
          rq[0] = rpre[RES_X && !HL ? 2 * ii : 0];
          rq[1] = rpre[RES_X && !HL ? 2 * ii + 1 : 0];
        }
        uint32_t acc[ECOLS];
        tmem_ld16(lane_taddr + b * (uint32_t)(MT * C) + (uint32_t)(m * C + cb * ECOLS), acc);
        if (ring_slot >= 0) {
          // The slot may be refilled (async proxy) as soon as r_empty completes, and an mbarrier arrive
          // does not wait for this warp's outstanding shared-memory LOADS: make the arrive data-dependent
          // on every loaded register so the scoreboard orders it after the loads.
          uint32_t dep = rq[0].x ^ rq[1].x ^ rq[2].x ^ rq[3].x ^ rq[0].w ^ rq[1].w ^ rq[2].w ^ rq[3].w;
          asm volatile("" : "+r"(dep));
          __syncwarp();
          if (lane == 0) mbar_arrive_dep(&r_empty[ring_slot], dep);
        }
        if (ok) {
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const size_t off = plane_base + (size_t)(cb * 2 + jj) * p.L + t;   // 8-element units
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[jj * 8 + i]) + bias2[cb * ECOLS + jj * 8 + i];
            if (!RES_X) {
              const uint4 q0 = rq[2 * jj], q1 = rq[2 * jj + 1];
              v[0] += __uint_as_float(q0.x); v[1] += __uint_as_float(q0.y);
              v[2] += __uint_as_float(q0.z); v[3] += __uint_as_float(q0.w);
              v[4] += __uint_as_float(q1.x); v[5] += __uint_as_float(q1.y);
              v[6] += __uint_as_float(q1.z); v[7] += __uint_as_float(q1.w);
            } else {
              float rr[8];
              unpack8(rq[jj], rr);
              if (HL) {
                float ll[8];
                unpack8(rq[2 + jj], ll);
#pragma unroll
                for (int i = 0; i < 8; ++i) rr[i] += ll[i];
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += fminf(rr[i], rr[i] * p.res_inv);
            }
            if (p.out_scale != 1.f) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] *= p.out_scale;
            }
            if (p.accin16) {
              float a[8];
              unpack8(pre16 ? pre[2 * ii + jj] : *(reinterpret_cast<const uint4*>(p.accin16) + off), a);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += a[i];
            }
            if (p.accin32) {
              uint4 a0, a1;
              if (PRE32 && pre32) {
                a0 = pre[(4 * ii + 2 * jj) & 7];
                a1 = pre[(4 * ii + 2 * jj + 1) & 7];
              } else {
                a0 = *(reinterpret_cast<const uint4*>(p.accin32) + off * 2);
                a1 = *(reinterpret_cast<const uint4*>(p.accin32) + off * 2 + 1);
              }
              v[0] += __uint_as_float(a0.x); v[1] += __uint_as_float(a0.y);
              v[2] += __uint_as_float(a0.z); v[3] += __uint_as_float(a0.w);
              v[4] += __uint_as_float(a1.x); v[5] += __uint_as_float(a1.y);
              v[6] += __uint_as_float(a1.z); v[7] += __uint_as_float(a1.w);
            }
            if (p.out32) {
              float4* o4 = reinterpret_cast<float4*>(p.out32) + off * 2;
              o4[0] = make_float4(v[0], v[1], v[2], v[3]);
              o4[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
            if (p.out16) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], v[i] * p.out16_slope);   // slope <= 1
              const uint4 hi = pack8(v);
              *(reinterpret_cast<uint4*>(p.out16) + off) = hi;
              if (HL && p.out_lo) {
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
      if (lane == 0) mbar_arrive(&acc2_empty[b]);
      ++j;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tcgen05_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

struct PairPlan {
  int MT, a_slots, a_slot_bytes, plane_bytes, tmp_plane_bytes, tmp_bytes, w_tile_bytes, r_slots, r_slot_bytes,
      tmem_cols, depth;
  size_t smem;
};

bool make_pair_plan(const PairConvArgs& a, PairPlan* out) {
  if (a.C != 32 && a.C != 64) return false;
  if (a.K < 1 || a.K > 16 || !(a.K & 1) || a.dil < 1) return false;
  const int planes = a.C / 8;
  const int w_tile = a.C * a.C * 2;
  const size_t fixed = 1024 + 1024 + 8 * (size_t)a.C + (a.post_part ? 4 * (size_t)K_POST * a.C : 0);   // alignment slack, barriers, two bias vectors, conv_post weights
  const size_t budget = (size_t)227 * 1024;
  const size_t w_all = 2 * (size_t)a.K * w_tile;
  static const int forced_mt = [] { const char* e = getenv("PG_PAIR_MT"); return e ? atoi(e) : 0; }();
  static const int forced_depth = [] { const char* e = getenv("PG_PAIR_DEPTH"); return e ? atoi(e) : 0; }();
  // MT: measured on the bench clip (profiles/r02f): C = 64 prefers 128-row tiles (k = 3: 0.148 vs 0.200 ms), C = 32
  // 256-row tiles (0.204 vs 0.245 / 0.230 ms at MT = 1 / 4).  pass 0 insists on rings deep enough to hide the HBM
  // latency (3 input windows in flight and, for the fp32 residual stream, two tiles of residual slots) and on a
  // pipeline depth of 3; pass 1 takes whatever fits.
  const int order64[3] = {1, 2, 0}, order32[3] = {2, 4, 1};
  for (int pass = 0; pass < 2; ++pass)
  for (int mi = 0; mi < 3; ++mi) {
    const int MT = a.C == 64 ? order64[mi] : order32[mi];
    if (MT == 0) continue;
    if (forced_mt && MT != forced_mt) continue;
    if (4 * MT * a.C > 512) continue;
    const int MO = BM * MT - (a.K - 1);
    if (MO < 64) continue;
    const int rows = BM * MT + (a.K - 1) * a.dil;
    const int plane_bytes = 16 * ((rows + 7) & ~7);
    const int a_slot = (planes * plane_bytes + 127) & ~127;
    const int tmp_plane = 16 * (BM * MT + 16);
    const int tmp_bytes = (planes * tmp_plane + 127) & ~127;
    const bool ring = a.res32 || a.x_lo;       // fp32 stream or hi/lo planes: [planes][128 rows] x 32 B per slot
    const int r_slot = ring ? BM * planes * 32 : 0;
    int r_slots = ring ? 2 : 0;
    size_t need = fixed + w_all + 2 * (size_t)a_slot + 2 * (size_t)tmp_bytes + (size_t)r_slots * r_slot;
    if (need > budget) continue;
    size_t left = budget - need;
    int a_slots = 2, depth = 2;
    // depth / lag sweeps (profiles/r02h: depth 2-4, lag 1-2) moved nothing: the stage that bounds the pairs is
    // not the GEMM 1 -> E1 -> GEMM 2 hand-off, so the shared memory goes to the rings instead
    const int max_depth = std::min(forced_depth > 0 ? forced_depth : 2, 512 / (2 * MT * a.C));
    auto grow = [&](int* v, int cap, size_t unit) {
      while (*v < cap && left >= unit) {
        ++*v;
        left -= unit;
      }
    };
    grow(&a_slots, 3, a_slot);
    if (ring) grow(&r_slots, 2 * MT, r_slot);
    grow(&depth, std::min(3, max_depth), tmp_bytes);
    grow(&a_slots, 4, a_slot);
    grow(&depth, max_depth, tmp_bytes);
    if (ring) grow(&r_slots, 3 * MT, r_slot);
    if (pass == 0 && (a_slots < 3 || (ring && r_slots < 2 * MT) || depth < std::min(3, max_depth))) continue;
    int tm = 32;
    while (tm < 2 * depth * MT * a.C) tm <<= 1;
    *out = PairPlan{MT, a_slots, a_slot, plane_bytes, tmp_plane, tmp_bytes, w_tile, r_slots, r_slot, tm, depth,
                    fixed + w_all + (size_t)a_slots * a_slot + (size_t)depth * tmp_bytes + (size_t)r_slots * r_slot};
    return true;
  }
  return false;
}

template <int MT, int C, bool RES_X, int EPIM, bool HL>
cudaError_t launch_pair_t(const PairConvArgs& a, const PairPlan& pl, cudaStream_t s) {
  CUtensorMap m1, m2;
  const int kc = C == 64 ? 64 : 32;
  if (!get_weight_map(a.w1, C, C, a.K, kc, C, &m1) || !get_weight_map(a.w2, C, C, a.K, kc, C, &m2))
    return cudaErrorNotSupported;
  PairParams p;
  p.x = a.x; p.L = a.L; p.B = a.B; p.K = a.K; p.dil = a.dil;
  p.post_w = a.post_w; p.post_part = a.post_part; p.post_slope = a.post_slope;
  p.tlen = a.tlen; p.len_mul = a.len_mul;
  p.bias1 = a.bias1; p.bias2 = a.bias2; p.res32 = a.res32; p.res_inv = a.res_inv;
  p.x_lo = a.x_lo; p.out_lo = a.out_lo;
  p.accin16 = a.accin16; p.accin32 = a.accin32;
  p.out16 = a.out16; p.out16_slope = a.out16_slope; p.out32 = a.out32; p.out_scale = a.out_scale;
  p.MO = BM * MT - (a.K - 1);
  p.n_row_tiles = (a.L + p.MO - 1) / p.MO;
  p.total_tiles = p.n_row_tiles * a.B;
  p.h1 = (a.K - 1) * a.dil / 2; p.h2 = (a.K - 1) / 2;
  p.plane_bytes = pl.plane_bytes; p.a_slots = pl.a_slots; p.a_slot_bytes = pl.a_slot_bytes;
  p.tmp_plane_bytes = pl.tmp_plane_bytes; p.tmp_bytes = pl.tmp_bytes; p.w_tile_bytes = pl.w_tile_bytes;
  p.r_slots = pl.r_slots; p.r_slot_bytes = pl.r_slot_bytes; p.tmem_cols = pl.tmem_cols; p.depth = pl.depth;
  static const int forced_lag = [] { const char* e = getenv("PG_PAIR_LAG"); return e ? atoi(e) : 0; }();
  p.lag = forced_lag > 0 ? std::min(forced_lag, pl.depth - 1) : pl.depth - 1;
  p.idesc = make_idesc(BM, C);
  static const int no_pre = [] { const char* e = getenv("PG_PAIR_NOPRE"); return e ? atoi(e) : 0; }();
  p.no_pre = no_pre;
  static const int no_ring = [] { const char* e = getenv("PG_PAIR_NORING"); return e ? atoi(e) : 0; }();
  p.no_ring = no_ring;

  static DeviceOnce once;
  if (cudaError_t e = ensure_dyn_smem(pair_planes_kernel<MT, C, RES_X, EPIM, HL>, once, 227 * 1024)) return e;
  int grid = device_sm_count();
  if (a.grid_cap > 0 && grid > a.grid_cap) grid = a.grid_cap;
  if (grid > p.total_tiles) grid = p.total_tiles;
  static const int dbg_sync = [] { const char* e = getenv("PG_PAIR_SYNC"); return e ? atoi(e) : 0; }();
  if (dbg_sync & 1) cudaStreamSynchronize(s);
  if (cudaError_t e = launch_pdl(pair_planes_kernel<MT, C, RES_X, EPIM, HL>, dim3(grid), dim3(NTHREADS), pl.smem, s, m1, m2, p)) return e;
  if (dbg_sync & 2) cudaStreamSynchronize(s);
  return cudaGetLastError();
}

}  // namespace

unsigned long long pair_debug_counter(int i) {
  unsigned long long v[16];
  cudaMemcpyFromSymbol(v, g_pair_dbg, sizeof(v));
  return v[i & 15];
}

bool pair_conv_supported(const PairConvArgs& a) {
  if (!a.x || !a.w1 || !a.w2 || !a.bias1 || !a.bias2 || (!a.out16 && !a.out32)) return false;
  if (a.L <= 0 || a.B <= 0 || a.res_inv < 1.f || a.out16_slope > 1.f) return false;
  if (a.x_lo && a.res32) return false;          // the residual is either hi + lo or the fp32 stream
  if (a.x_lo && a.C != 32) return false;        // the hi/lo ring loader serves 2 x 4 planes with 8 lanes
  if (a.out_lo && (!a.out16 || !a.x_lo)) return false;
  PairPlan pl;
  return make_pair_plan(a, &pl);
}

int pair_conv_mt(const PairConvArgs& a) {
  PairPlan pl;
  return make_pair_plan(a, &pl) ? pl.MT : 0;
}

cudaError_t launch_pair_planes(const PairConvArgs& a, cudaStream_t s) {
  PairPlan pl;
  if (!pair_conv_supported(a) || !make_pair_plan(a, &pl)) return cudaErrorInvalidValue;
  const bool hl = a.x_lo != nullptr;
  // straight-line E2 forms (see EPIM): 1 plain pair, 2 / 3 last pair of a ResBlock without / with the running mean
  int epim = 0;
  if (!a.res32) {
    if (hl) {
      if (!a.accin16 && !a.accin32 && !a.out32 && a.out16 && a.out_lo && a.out_scale == 1.f) epim = 1;
      else if (a.out32 && !a.out16 && !a.accin16) epim = a.accin32 ? 3 : 2;
    } else {
      if (!a.accin16 && !a.accin32 && !a.out32 && a.out16 && a.out_scale == 1.f) epim = 1;
      else if (a.out16 && !a.out32 && !a.accin32) epim = a.accin16 ? 3 : 2;
    }
  }
  static const int no_epim = [] { const char* e = getenv("PG_PAIR_GENERIC_E2"); return e ? atoi(e) : 0; }();
  if (no_epim) epim = 0;
  static const int epim_mask = [] { const char* e = getenv("PG_PAIR_EPIM_MASK"); return e ? atoi(e) : 14; }();
  if (!((epim_mask >> epim) & 1)) epim = 0;
  if (a.post_part && !(hl && epim >= 2 && a.post_w && a.post_slope <= 1.f)) return cudaErrorInvalidValue;   // only the straight-line last-pair E2 computes the conv_post partials
#define PG_PAIR_E(MT_, C_, HL_)                                                      \
  switch (epim) {                                                                   \
    case 1: return launch_pair_t<MT_, C_, true, 1, HL_>(a, pl, s);                   \
    case 2: return launch_pair_t<MT_, C_, true, 2, HL_>(a, pl, s);                   \
    case 3: return launch_pair_t<MT_, C_, true, 3, HL_>(a, pl, s);                   \
    default: return launch_pair_t<MT_, C_, true, 0, HL_>(a, pl, s);                  \
  }
#define PG_PAIR(MT_, C_)                                                             \
  do {                                                                              \
    if (a.res32) return launch_pair_t<MT_, C_, false, 0, false>(a, pl, s);           \
    if (hl) { PG_PAIR_E(MT_, C_, true) }                                             \
    PG_PAIR_E(MT_, C_, false)                                                        \
  } while (0)
  if (a.C == 32) {
    switch (pl.MT) {
      case 4: PG_PAIR(4, 32);
      case 2: PG_PAIR(2, 32);
      case 1: PG_PAIR(1, 32);
    }
  } else {
    switch (pl.MT) {
      case 2: PG_PAIR(2, 64);
      case 1: PG_PAIR(1, 64);
    }
  }
#undef PG_PAIR
#undef PG_PAIR_E
  return cudaErrorInvalidValue;
}

}  // namespace pg
