// Windowed relative-position multi-head attention of the TextEncoder on tensor cores
// (reference rvc/lib/algorithm/attentions.py:63-113, helpers :115-158):
//   scores[i,j] = q_i.k_j/sqrt(d) + [|j-i|<=w] q_i.Ek[j-i+w]/sqrt(d)
//   masked_fill(mask_i*mask_j == 0, -1e4); softmax_j
//   out_i = sum_j p_ij v_j + sum_{|j-i|<=w} p_ij Ev[j-i+w]
// 1.5 % of the path's FLOPs, but O(T^2): flash-style, 64 queries x 64 keys per step with online
// softmax, both GEMMs on mma.sync m16n8k16 (f16 operands, fp32 accumulate).  The latents this
// feeds are returned to the caller and drive the whole decoder, so every operand is split into
// two f16 terms (x = hi + lo) and each product is evaluated as hi*hi + hi*lo + lo*hi: ~2^-22
// relative error, i.e. fp32-class results at tensor-core speed.
// The relative-position terms are tiny (21 taps) and stay on CUDA cores in fp32.
// A segment has only T/64 * heads query tiles (38 for a 12 s segment), far fewer than the 148 SMs,
// and one warp's MMA chain over all T keys is latency-bound, so the key range is split over
// `splits` CTAs per query tile (flash-decoding style): each writes its unnormalised partial output,
// running max and sum plus its band scores, and attn_merge_kernel combines them, normalises and
// adds the relative-value term.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "pg_common.cuh"
#include "pg_umma.cuh"

namespace pg {

namespace {

constexpr int AQ = 64, AK = 64, AD = 96;   // queries / keys per step, head dim
constexpr int KP = 104;                    // K tile row pitch (halves): conflict-free B-fragment loads
constexpr int VP = 72;                     // V^T tile row pitch (halves)
constexpr int RP = 24;                     // rel-pos table pitch (floats), 2*window+1 <= RP
constexpr int ATT_THREADS = 128;

__device__ __forceinline__ void split_f16(float x, __half* hi, __half* lo) {
  const __half h = __float2half_rn(x);
  *hi = h;
  *lo = __float2half_rn(x - __half2float(h));
}

// qkv [B][T][3H] f32 -> Q (pre-scaled), K as [B][heads][Tp][96] hi/lo and V^T as [B][heads][96][Tp]
// hi/lo; rows >= T are zero.
__global__ void attn_prep_kernel(const float* __restrict__ qkv, __half* __restrict__ qh, __half* __restrict__ ql,
                                 __half* __restrict__ kh, __half* __restrict__ kl, __half* __restrict__ vth,
                                 __half* __restrict__ vtl, int B, int T, int Tp, int H, int heads, float qscale) {
  const size_t total = (size_t)B * heads * Tp * AD;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % AD);
  const size_t r = i / AD;
  const int j = (int)(r % Tp);
  const size_t bh = r / Tp;
  const int h = (int)(bh % heads), b = (int)(bh / heads);
  float q = 0.f, k = 0.f, v = 0.f;
  if (j < T) {
    const float* row = qkv + ((size_t)b * T + j) * 3 * H + h * AD + c;
    q = row[0] * qscale;
    k = row[H];
    v = row[2 * H];
  }
  split_f16(q, qh + i, ql + i);
  split_f16(k, kh + i, kl + i);
  const size_t vt = (bh * AD + c) * Tp + j;
  split_f16(v, vth + vt, vtl + vt);
}

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  const __half2 h = __halves2half2(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

struct AttnSmem {
  __half k[1][2][AK][KP];    // [buffer][hi/lo][key][dim]  (single buffer: two CTAs per SM overlap instead)
  __half v[1][2][AD][VP];    // [buffer][hi/lo][dim][key]
  float rl[AQ][RP];          // rel-key logits q_i . Ek[r]
  float qs[AQ][AD + 1];      // fp32 query rows / rel-key table staged for the rl prologue
  float ek[RP][AD + 1];      //   (odd pitch: conflict-free for the row-per-thread reads)
};

__global__ void __launch_bounds__(ATT_THREADS, 2) rel_attention_mma_kernel(
    const __half* __restrict__ qh, const __half* __restrict__ ql, const __half* __restrict__ kh,
    const __half* __restrict__ kl, const __half* __restrict__ vth, const __half* __restrict__ vtl,
    const float* __restrict__ qkv, const float* __restrict__ rel_k, const int* __restrict__ lens,
    float* __restrict__ part_o, float* __restrict__ part_m, float* __restrict__ part_l,
    float* __restrict__ band_s, int B, int T, int Tp, int H, int heads, int window, int splits, float qscale) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  AttnSmem& sm = *reinterpret_cast<AttnSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z / splits, sp = blockIdx.z - b * splits, h = blockIdx.y, q0 = blockIdx.x * AQ;
  const int len = lens[b];
  // Rows of a (ragged / bucket-padded) batch end at len: query tiles past it produce nothing anybody reads (every
  // consumer masks rows >= len; the merge kernel writes zeros there), and key blocks past it carry exactly zero
  // probability for the live queries (exp(-1e4 - max) underflows), so both are skipped.
  if (q0 >= len) return;
  const int R = 2 * window + 1;
  const size_t bh = (size_t)b * heads + h;
  const __half* kh_b = kh + bh * Tp * AD;
  const __half* kl_b = kl + bh * Tp * AD;
  const __half* vth_b = vth + bh * AD * Tp;
  const __half* vtl_b = vtl + bh * AD * Tp;

  auto load_block = [&](int kb, int buf) {
    const int k0 = kb * AK;
    // K: 64 rows x 12 chunks of 16 B, hi and lo
    for (int i = tid; i < AK * 12; i += ATT_THREADS) {
      const int r = i / 12, ch = i - r * 12;
      cp_async16(&sm.k[buf][0][r][ch * 8], kh_b + (size_t)(k0 + r) * AD + ch * 8);
      cp_async16(&sm.k[buf][1][r][ch * 8], kl_b + (size_t)(k0 + r) * AD + ch * 8);
    }
    // V^T: 96 rows x 8 chunks
    for (int i = tid; i < AD * 8; i += ATT_THREADS) {
      const int r = i >> 3, ch = i & 7;
      cp_async16(&sm.v[buf][0][r][ch * 8], vth_b + (size_t)r * Tp + k0 + ch * 8);
      cp_async16(&sm.v[buf][1][r][ch * 8], vtl_b + (size_t)r * Tp + k0 + ch * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int nblk = Tp / AK;
  const int kb_begin = (int)((long long)sp * nblk / splits);
  const int kb_end = min((int)((long long)(sp + 1) * nblk / splits), (len + AK - 1) / AK);
  if (kb_begin >= kb_end) {     // this split holds no live key: a neutral partial (weight e^{-inf} = 0 in the merge)
    const size_t rows_all = (size_t)B * heads * Tp;
    for (int i = tid; i < AQ; i += ATT_THREADS) {
      part_m[(size_t)sp * rows_all + bh * Tp + q0 + i] = -INFINITY;
      part_l[(size_t)sp * rows_all + bh * Tp + q0 + i] = 0.f;
    }
    for (int i = tid; i < AQ * AD; i += ATT_THREADS)
      part_o[((size_t)sp * rows_all + bh * Tp + q0) * AD + i] = 0.f;
    return;
  }
  load_block(kb_begin, 0);

  // ---- prologue: rel-key logits (fp32) if this CTA's keys touch the band of its queries, Q fragments
  if (kb_begin * AK < q0 + AQ + window && kb_end * AK > q0 - window) {
    for (int i = tid; i < AQ * (AD / 4); i += ATT_THREADS) {
      const int il = i / (AD / 4), c4 = i - il * (AD / 4);
      const int qi = q0 + il;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (qi < T) v = *reinterpret_cast<const float4*>(qkv + ((size_t)b * T + qi) * 3 * H + h * AD + c4 * 4);
      sm.qs[il][c4 * 4 + 0] = v.x * qscale;
      sm.qs[il][c4 * 4 + 1] = v.y * qscale;
      sm.qs[il][c4 * 4 + 2] = v.z * qscale;
      sm.qs[il][c4 * 4 + 3] = v.w * qscale;
    }
    for (int i = tid; i < R * AD; i += ATT_THREADS) {
      const int r = i / AD;
      sm.ek[r][i - r * AD] = rel_k[i];
    }
    __syncthreads();
    for (int i = tid; i < AQ * RP; i += ATT_THREADS) {
      const int il = i / RP, r = i - il * RP;
      float s = 0.f;
      if (r < R) {
#pragma unroll 8
        for (int c = 0; c < AD; ++c) s = fmaf(sm.qs[il][c], sm.ek[r][c], s);
      }
      sm.rl[il][r] = s;
    }
  }
  const int r0 = q0 + warp * 16;            // this warp's 16 query rows
  const int i0 = r0 + g, i1 = r0 + g + 8;   // this lane's two rows
  float* band0 = band_s + (bh * Tp + i0) * RP;
  float* band1 = band_s + (bh * Tp + i1) * RP;
  uint32_t aqh[6][4], aql[6][4];
  {
    const __half* qh_b = qh + bh * Tp * AD;
    const __half* ql_b = ql + bh * Tp * AD;
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) {
      const int c = ks * 16 + 2 * t;
      aqh[ks][0] = *reinterpret_cast<const uint32_t*>(qh_b + (size_t)i0 * AD + c);
      aqh[ks][1] = *reinterpret_cast<const uint32_t*>(qh_b + (size_t)i1 * AD + c);
      aqh[ks][2] = *reinterpret_cast<const uint32_t*>(qh_b + (size_t)i0 * AD + c + 8);
      aqh[ks][3] = *reinterpret_cast<const uint32_t*>(qh_b + (size_t)i1 * AD + c + 8);
      aql[ks][0] = *reinterpret_cast<const uint32_t*>(ql_b + (size_t)i0 * AD + c);
      aql[ks][1] = *reinterpret_cast<const uint32_t*>(ql_b + (size_t)i1 * AD + c);
      aql[ks][2] = *reinterpret_cast<const uint32_t*>(ql_b + (size_t)i0 * AD + c + 8);
      aql[ks][3] = *reinterpret_cast<const uint32_t*>(ql_b + (size_t)i1 * AD + c + 8);
    }
  }
  float o[12][4];
#pragma unroll
  for (int n = 0; n < 12; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[n][e] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int kb = kb_begin; kb < kb_end; ++kb) {
    const int buf = 0, k0 = kb * AK;
    if (kb > kb_begin) load_block(kb, 0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- S = Q K^T (3-term split)
    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[n][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) {
      // independent accumulators back to back: the three split terms of one tile are issued
      // 8 MMAs apart so no MMA waits on its predecessor
      uint32_t bh0[8], bh1[8], bl0[8], bl1[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const __half* kr_h = &sm.k[buf][0][n * 8 + g][ks * 16 + 2 * t];
        const __half* kr_l = &sm.k[buf][1][n * 8 + g][ks * 16 + 2 * t];
        bh0[n] = *reinterpret_cast<const uint32_t*>(kr_h);
        bh1[n] = *reinterpret_cast<const uint32_t*>(kr_h + 8);
        bl0[n] = *reinterpret_cast<const uint32_t*>(kr_l);
        bl1[n] = *reinterpret_cast<const uint32_t*>(kr_l + 8);
      }
#pragma unroll
      for (int n = 0; n < 8; ++n) mma16816(s[n], aqh[ks], bh0[n], bh1[n]);
#pragma unroll
      for (int n = 0; n < 8; ++n) mma16816(s[n], aqh[ks], bl0[n], bl1[n]);
#pragma unroll
      for (int n = 0; n < 8; ++n) mma16816(s[n], aql[ks], bh0[n], bh1[n]);
    }
    // ---- rel-key logits on the band, masks
    const bool band = (k0 < q0 + AQ + window) && (k0 + AK > q0 - window);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = k0 + n * 8 + 2 * t + (e & 1);
        const int i = e < 2 ? i0 : i1;
        float v = s[n][e];
        const int rel = j - i + window;
        const bool inband = band && rel >= 0 && rel < R;
        if (inband) v += sm.rl[i - q0][rel];
        if (!(i < len && j < len)) v = -1e4f;
        if (j >= T) v = -INFINITY;
        if (inband && j < T) (e < 2 ? band0 : band1)[rel] = v;   // each (i, j) of the band has one owner CTA
        s[n][e] = v;
        if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float al0 = __expf(m0 - mn0), al1 = __expf(m1 - mn1);
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      s[n][0] = __expf(s[n][0] - mn0);
      s[n][1] = __expf(s[n][1] - mn0);
      s[n][2] = __expf(s[n][2] - mn1);
      s[n][3] = __expf(s[n][3] - mn1);
      rs0 += s[n][0] + s[n][1];
      rs1 += s[n][2] + s[n][3];
    }
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1);
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
    l0 = l0 * al0 + rs0;
    l1 = l1 * al1 + rs1;
    m0 = mn0;
    m1 = mn1;
#pragma unroll
    for (int n = 0; n < 12; ++n) {
      o[n][0] *= al0; o[n][1] *= al0;
      o[n][2] *= al1; o[n][3] *= al1;
    }
    // ---- O += P V (3-term split; P fragments come straight from the S accumulators)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int n = 2 * kk + half;
        __half h0, h1, h2, h3, e0, e1, e2, e3;
        split_f16(s[n][0], &h0, &e0);
        split_f16(s[n][1], &h1, &e1);
        split_f16(s[n][2], &h2, &e2);
        split_f16(s[n][3], &h3, &e3);
        ph[2 * half] = pack_h2(h0, h1);
        ph[2 * half + 1] = pack_h2(h2, h3);
        pl[2 * half] = pack_h2(e0, e1);
        pl[2 * half + 1] = pack_h2(e2, e3);
      }
#pragma unroll
      for (int n6 = 0; n6 < 12; n6 += 6) {
        uint32_t bh0[6], bh1[6], bl0[6], bl1[6];
#pragma unroll
        for (int n = 0; n < 6; ++n) {
          const __half* vr_h = &sm.v[buf][0][(n6 + n) * 8 + g][kk * 16 + 2 * t];
          const __half* vr_l = &sm.v[buf][1][(n6 + n) * 8 + g][kk * 16 + 2 * t];
          bh0[n] = *reinterpret_cast<const uint32_t*>(vr_h);
          bh1[n] = *reinterpret_cast<const uint32_t*>(vr_h + 8);
          bl0[n] = *reinterpret_cast<const uint32_t*>(vr_l);
          bl1[n] = *reinterpret_cast<const uint32_t*>(vr_l + 8);
        }
#pragma unroll
        for (int n = 0; n < 6; ++n) mma16816(o[n6 + n], ph, bh0[n], bh1[n]);
#pragma unroll
        for (int n = 0; n < 6; ++n) mma16816(o[n6 + n], ph, bl0[n], bl1[n]);
#pragma unroll
        for (int n = 0; n < 6; ++n) mma16816(o[n6 + n], pl, bh0[n], bh1[n]);
      }
    }
    __syncthreads();   // this buffer is refilled by the next iteration's prefetch
  }

  // ---- epilogue: this split's running max / sum and unnormalised output
  const size_t rows = (size_t)B * heads * Tp;
  const size_t row0 = (size_t)sp * rows + bh * Tp + i0, row1 = row0 + 8;
  if (t == 0) {
    part_m[row0] = m0;
    part_l[row0] = l0;
    part_m[row1] = m1;
    part_l[row1] = l1;
  }
#pragma unroll
  for (int n = 0; n < 12; ++n) {
    const int c = n * 8 + 2 * t;
    *reinterpret_cast<float2*>(part_o + row0 * AD + c) = make_float2(o[n][0], o[n][1]);
    *reinterpret_cast<float2*>(part_o + row1 * AD + c) = make_float2(o[n][2], o[n][3]);
  }
}

// out_i = (sum_s O_s e^{m_s-M}) / L + sum_r p_{i,i+r-w} Ev[r],  M = max_s m_s,  L = sum_s l_s e^{m_s-M}.
// One warp per (b, head, query row); lanes cover the 96 channels three at a time.
__global__ void attn_merge_kernel(const float* __restrict__ part_o, const float* __restrict__ part_m,
                                  const float* __restrict__ part_l, const float* __restrict__ band_s,
                                  const float* __restrict__ rel_v, const int* __restrict__ lens,
                                  float* __restrict__ out, int B, int T, int Tp, int H, int heads, int window,
                                  int splits) {
  const int lane = threadIdx.x & 31;
  const size_t gw = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= (size_t)B * heads * T) return;
  const int i = (int)(gw % T);
  const size_t bh = gw / T;
  const int h = (int)(bh % heads), b = (int)(bh / heads);
  const size_t rows = (size_t)B * heads * Tp, row = bh * Tp + i;
  const int len = lens[b];
  if (i >= len) {      // past the row's end: no partials were produced; consumers mask these rows anyway
    float* dst0 = out + ((size_t)b * T + i) * H + h * AD + lane;
#pragma unroll
    for (int k = 0; k < 3; ++k) dst0[32 * k] = 0.f;
    return;
  }
  float M = -INFINITY;
  for (int s = 0; s < splits; ++s) M = fmaxf(M, part_m[s * rows + row]);
  float L = 0.f, acc[3] = {0.f, 0.f, 0.f};
  for (int s = 0; s < splits; ++s) {
    const float w = __expf(part_m[s * rows + row] - M);
    L = fmaf(part_l[s * rows + row], w, L);
    const float* po = part_o + (s * rows + row) * AD + lane;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[k] = fmaf(po[32 * k], w, acc[k]);
  }
  const float inv_l = 1.f / L;
  const int R = 2 * window + 1;
  float p = 0.f;   // band probability of key j = i + lane - window
  if (lane < R) {
    const int j = i + lane - window;
    if (j >= 0 && j < len) p = __expf(band_s[row * RP + lane] - M) * inv_l;   // keys >= len: probability exactly 0
  }
  float v[3] = {acc[0] * inv_l, acc[1] * inv_l, acc[2] * inv_l};
  for (int r = 0; r < R; ++r) {
    const float pr = __shfl_sync(0xffffffffu, p, r);
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = fmaf(pr, rel_v[r * AD + lane + 32 * k], v[k]);
  }
  float* dst = out + ((size_t)b * T + i) * H + h * AD + lane;
#pragma unroll
  for (int k = 0; k < 3; ++k) dst[32 * k] = v[k];
}

// =====================================================================================================
// tcgen05 / TMEM version (the product path).  Same algorithm and the same split-key partials + merge kernel as
// the mma.sync kernel above (kept as the validation twin, PG_ATTN_LEGACY=1), but both GEMMs run on the 5th-gen
// tensor cores with fp32 accumulators in TMEM:
//   * one CTA = 128 queries x a range of 64-key blocks; 6 warps: 4 softmax warps (thread = query row =
//     TMEM lane), 1 MMA-issue warp, 1 bulk-copy loader warp
//   * S = Q K^T: M = 128, N = 64, K = 96 as 3 split terms x 6 MMAs (Qh Kh + Qh Kl + Ql Kh) into one of two
//     S accumulators (64 TMEM columns each); S(j+2) is issued right after PV(j), so the tensor pipe works on
//     the next blocks' scores while the softmax warps are busy with block j+1
//   * P = exp(S - m) leaves the softmax threads as a two-term f16 pair in shared memory in the canonical
//     no-swizzle K-major operand layout ([key group of 8][row][8 keys]: a thread's 16 B stores are conflict-free)
//   * O_j = P V: M = 128, N = 96, K = 64 as 3 terms x 4 MMAs into one of two O accumulators (96 columns, fresh
//     per block); the running output stays in REGISTERS (o = o*alpha + O_j after a tcgen05.ld), so the online-
//     softmax rescale never touches TMEM and the read of O_{j-1} is deferred behind the softmax of block j
//   * operands come pre-split and pre-tiled from attn_prep_umma_kernel: a Q tile (48 KB), a K block and a V
//     block (24 KB each; V as [key group][dim][8 keys]) are each ONE contiguous bulk copy (cp.async.bulk)
//     completing on an mbarrier; K has a 2-deep ring (released by the commit of its S MMAs), V a 3-deep one
//     (released by the commit of its PV MMAs)
// =====================================================================================================
constexpr int UQ = 128, UK = 64;                    // queries per CTA, keys per block
constexpr int U_PLANES = AD / 8;                    // 12 planes of 8 head dims
constexpr int UQ_TERM = U_PLANES * UQ * 16;         // 24576 B: one term (hi | lo) of a Q tile
constexpr int UK_TERM = U_PLANES * UK * 16;         // 12288 B: one term of a K block  [plane][key][8]
constexpr int UV_TERM = (UK / 8) * AD * 16;         // 12288 B: one term of a V block  [key group][dim][8 keys]
constexpr int UP_TERM = (UK / 8) * UQ * 16;         // 16384 B: one term of P          [key group][row][8 keys]
constexpr int U_KSTAGES = 2, U_VSTAGES = 3;
constexpr int URL = 22;                             // rel-logit row pitch (floats): 21*i + j + w is conflict-free
constexpr int U_OFF_Q = 0;
constexpr int U_OFF_K = U_OFF_Q + 2 * UQ_TERM;
constexpr int U_OFF_V = U_OFF_K + U_KSTAGES * 2 * UK_TERM;
constexpr int U_OFF_P = U_OFF_V + U_VSTAGES * 2 * UV_TERM;      // prologue: aliased by the fp32 rel-key table
constexpr int U_OFF_RL = U_OFF_P + 2 * UP_TERM;
constexpr int U_OFF_BAR = U_OFF_RL + UQ * URL * 4;
constexpr int U_OFF_XM = U_OFF_BAR + 256;           // NH = 2: row-max exchange between the two column halves [2][2][128] f32
constexpr int U_SMEM = U_OFF_XM + 2 * 2 * UQ * 4;
constexpr uint32_t U_TMEM_COLS = 512;               // S0 @0, S1 @64, O0 @128, O1 @256
static_assert(RP * AD * 4 <= 2 * UP_TERM, "rel-key table must fit the P buffer it aliases");
static_assert(U_SMEM <= 227 * 1024, "attention tile set exceeds shared memory");

// qkv [B][T][3H] f32 -> pre-scaled Q tiles, K blocks, V blocks (two f16 terms each) in the layouts above; rows >= T
// are zero.  One CTA per (b, head, 64-row block): items are (row, plane) for Q / K and (key group, dim) for V, so
// the 16 B stores of consecutive threads are contiguous.
__global__ void attn_prep_umma_kernel(const float* __restrict__ qkv, __half* __restrict__ qb, __half* __restrict__ kb,
                                      __half* __restrict__ vb, int T, int Tp, int H, int heads, float qscale) {
  const int blk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const size_t bh = (size_t)b * heads + h;
  const int j0 = blk * UK;
  char* q_tile = reinterpret_cast<char*>(qb) + (bh * (Tp / UQ) + (size_t)(j0 / UQ)) * (2 * UQ_TERM);
  char* k_blk = reinterpret_cast<char*>(kb) + (bh * (Tp / UK) + (size_t)blk) * (2 * UK_TERM);
  char* v_blk = reinterpret_cast<char*>(vb) + (bh * (Tp / UK) + (size_t)blk) * (2 * UV_TERM);
  const int rq0 = j0 % UQ;      // first row of this block inside its Q tile
  for (int it = threadIdx.x; it < UK * U_PLANES; it += blockDim.x) {
    const int r = it % UK, pl = it / UK;
    const int j = j0 + r;
    float q[8], k[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) q[e] = k[e] = 0.f;
    if (j < T) {
      const float* row = qkv + ((size_t)b * T + j) * 3 * H + h * AD + pl * 8;
      const float4 q0 = *reinterpret_cast<const float4*>(row), q1 = *reinterpret_cast<const float4*>(row + 4);
      const float4 k0 = *reinterpret_cast<const float4*>(row + H), k1 = *reinterpret_cast<const float4*>(row + H + 4);
      q[0] = q0.x * qscale; q[1] = q0.y * qscale; q[2] = q0.z * qscale; q[3] = q0.w * qscale;
      q[4] = q1.x * qscale; q[5] = q1.y * qscale; q[6] = q1.z * qscale; q[7] = q1.w * qscale;
      k[0] = k0.x; k[1] = k0.y; k[2] = k0.z; k[3] = k0.w; k[4] = k1.x; k[5] = k1.y; k[6] = k1.z; k[7] = k1.w;
    }
    __align__(16) __half qh[8], ql[8], kh[8], kl[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      split_f16(q[e], &qh[e], &ql[e]);
      split_f16(k[e], &kh[e], &kl[e]);
    }
    const size_t qo = ((size_t)pl * UQ + rq0 + r) * 16, ko = ((size_t)pl * UK + r) * 16;
    *reinterpret_cast<uint4*>(q_tile + qo) = *reinterpret_cast<const uint4*>(qh);
    *reinterpret_cast<uint4*>(q_tile + UQ_TERM + qo) = *reinterpret_cast<const uint4*>(ql);
    *reinterpret_cast<uint4*>(k_blk + ko) = *reinterpret_cast<const uint4*>(kh);
    *reinterpret_cast<uint4*>(k_blk + UK_TERM + ko) = *reinterpret_cast<const uint4*>(kl);
  }
  for (int it = threadIdx.x; it < (UK / 8) * AD; it += blockDim.x) {
    const int c = it % AD, kg = it / AD;
    __align__(16) __half vh[8], vl[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = j0 + kg * 8 + e;
      const float v = j < T ? qkv[((size_t)b * T + j) * 3 * H + 2 * H + h * AD + c] : 0.f;
      split_f16(v, &vh[e], &vl[e]);
    }
    const size_t vo = ((size_t)kg * AD + c) * 16;
    *reinterpret_cast<uint4*>(v_blk + vo) = *reinterpret_cast<const uint4*>(vh);
    *reinterpret_cast<uint4*>(v_blk + UV_TERM + vo) = *reinterpret_cast<const uint4*>(vl);
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// NH = softmax threads per query row.  NH = 1: four softmax warps, a thread owns its row's 64 scores and 96 outputs.
// NH = 2: eight softmax warps -- warps w and w + 4 share TMEM lane quarter w and split the row's columns (32 scores,
// 48 outputs each); the only per-block exchange is the row maximum (one float through shared memory + a 64-thread named
// barrier), the row sums stay per-thread partials until the end.  Two warps per scheduler hide the latency one warp
// exposes (ncu on NH = 1: 4.7 cycles per issued instruction), and a thread's 80 live values fit the 168 registers
// ten warps leave it.
template <int NH>
__global__ void __launch_bounds__(128 * NH + 64, 1) rel_attention_umma_kernel(
    const __half* __restrict__ qb, const __half* __restrict__ kb, const __half* __restrict__ vb,
    const float* __restrict__ qkv, const float* __restrict__ rel_k, const int* __restrict__ lens,
    float* __restrict__ part_o, float* __restrict__ part_m, float* __restrict__ part_l, float* __restrict__ band_s,
    int B, int T, int Tp, int H, int heads, int window, int splits, float qscale, uint32_t idesc_s,
    uint32_t idesc_o) {
  using namespace umma;
  constexpr int U_THREADS = 128 * NH + 64, U_SOFT = 128 * NH, MMA_W = 4 * NH, LOAD_W = 4 * NH + 1;
  constexpr int NC = UK / NH, NO = AD / NH;     // score columns / output columns per softmax thread
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.z / splits, sp = blockIdx.z - b * splits, h = blockIdx.y, q0 = blockIdx.x * UQ;
  const int len = lens[b];
  if (q0 >= len) return;       // see the mma.sync kernel: nothing reads these rows
  const int R = 2 * window + 1;
  const size_t bh = (size_t)b * heads + h;
  const size_t rows_all = (size_t)B * heads * Tp;
  const int nblk = Tp / UK;
  const int kb_begin = (int)((long long)sp * nblk / splits);
  const int kb_end = min((int)((long long)(sp + 1) * nblk / splits), (len + UK - 1) / UK);
  if (kb_begin >= kb_end) {    // no live key in this split: a neutral partial
    for (int i = tid; i < UQ; i += U_THREADS) {
      part_m[(size_t)sp * rows_all + bh * Tp + q0 + i] = -INFINITY;
      part_l[(size_t)sp * rows_all + bh * Tp + q0 + i] = 0.f;
    }
    for (int i = tid; i < UQ * AD; i += U_THREADS) part_o[((size_t)sp * rows_all + bh * Tp + q0) * AD + i] = 0.f;
    return;
  }
  const int n = kb_end - kb_begin;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + U_OFF_BAR);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // 2
  uint64_t* k_empty = bars + 3;            // 2
  uint64_t* v_full = bars + 5;             // 3
  uint64_t* v_empty = bars + 8;            // 3
  uint64_t* s_full = bars + 11;            // 2
  uint64_t* o_full = bars + 13;            // 2
  uint64_t* p_full = bars + 15;            // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  if (tid == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&o_full[i], 1); }
    mbar_init(p_full, U_SOFT);
    fence_barrier_init();
  }
  if (warp == MMA_W) tcgen05_alloc(tmem_slot, U_TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == LOAD_W) {
    // ------------------------------------------------ loader: one bulk copy per Q tile / K block / V block
    if (elect_one()) {
      const char* q_src = reinterpret_cast<const char*>(qb) + (bh * (Tp / UQ) + (size_t)blockIdx.x) * (2 * UQ_TERM);
      mbar_expect_tx(q_full, 2 * UQ_TERM);
      bulk_g2s(smem + U_OFF_Q, q_src, 2 * UQ_TERM, q_full);
      const char* k_src = reinterpret_cast<const char*>(kb) + bh * nblk * (size_t)(2 * UK_TERM);
      const char* v_src = reinterpret_cast<const char*>(vb) + bh * nblk * (size_t)(2 * UV_TERM);
      for (int jj = 0; jj < n; ++jj) {
        const int ks = jj % U_KSTAGES, vs = jj % U_VSTAGES;
        mbar_wait(&k_empty[ks], ((jj / U_KSTAGES) & 1) ^ 1);
        mbar_expect_tx(&k_full[ks], 2 * UK_TERM);
        bulk_g2s(smem + U_OFF_K + ks * 2 * UK_TERM, k_src + (size_t)(kb_begin + jj) * (2 * UK_TERM), 2 * UK_TERM,
                 &k_full[ks]);
        mbar_wait(&v_empty[vs], ((jj / U_VSTAGES) & 1) ^ 1);
        mbar_expect_tx(&v_full[vs], 2 * UV_TERM);
        bulk_g2s(smem + U_OFF_V + vs * 2 * UV_TERM, v_src + (size_t)(kb_begin + jj) * (2 * UV_TERM), 2 * UV_TERM,
                 &v_full[vs]);
      }
    }
  } else if (warp == MMA_W) {
    // ------------------------------------------------ MMA issue
    if (elect_one()) {
      const uint32_t q_s = smem_u32(smem + U_OFF_Q), k_s = smem_u32(smem + U_OFF_K), v_s = smem_u32(smem + U_OFF_V),
                     p_s = smem_u32(smem + U_OFF_P);
      auto issue_s = [&](int jj) {      // S(jj) = Qh Kh^T + Qh Kl^T + Ql Kh^T  ->  TMEM columns (jj & 1) * 64
        const int ks = jj % U_KSTAGES;
        mbar_wait(&k_full[ks], (jj / U_KSTAGES) & 1);
        tcgen05_fence_after();
        const uint32_t kbase = k_s + ks * 2 * UK_TERM, d = tmem + (uint32_t)(jj & 1) * UK;
        uint32_t acc = 0;
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t qa = q_s + (term == 2 ? UQ_TERM : 0), ka = kbase + (term == 1 ? UK_TERM : 0);
#pragma unroll
          for (int k16 = 0; k16 < AD / 16; ++k16) {
            umma_f16(d, make_desc(qa + k16 * 2 * (UQ * 16), UQ * 16, 128, LAYOUT_NONE),
                     make_desc(ka + k16 * 2 * (UK * 16), UK * 16, 128, LAYOUT_NONE), idesc_s, acc);
            acc = 1;
          }
        }
        tcgen05_commit(&s_full[jj & 1]);
        tcgen05_commit(&k_empty[ks]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      if (n > 1) issue_s(1);
      for (int jj = 0; jj < n; ++jj) {
        const int vs = jj % U_VSTAGES;
        mbar_wait(p_full, jj & 1);
        mbar_wait(&v_full[vs], (jj / U_VSTAGES) & 1);
        tcgen05_fence_after();
        // O(jj) = Ph Vh + Ph Vl + Pl Vh  ->  TMEM columns 128 + (jj & 1) * 128
        const uint32_t vbase = v_s + vs * 2 * UV_TERM, d = tmem + 128u + (uint32_t)(jj & 1) * 128u;
        uint32_t acc = 0;
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t pa = p_s + (term == 2 ? UP_TERM : 0), va = vbase + (term == 1 ? UV_TERM : 0);
#pragma unroll
          for (int k16 = 0; k16 < UK / 16; ++k16) {
            umma_f16(d, make_desc(pa + k16 * 2 * (UQ * 16), UQ * 16, 128, LAYOUT_NONE),
                     make_desc(va + k16 * 2 * (AD * 16), AD * 16, 128, LAYOUT_NONE), idesc_o, acc);
            acc = 1;
          }
        }
        tcgen05_commit(&o_full[jj & 1]);
        tcgen05_commit(&v_empty[vs]);
        if (jj + 2 < n) issue_s(jj + 2);
      }
    }
  } else {
    // ------------------------------------------------ softmax: thread = (query row = TMEM lane, column half)
    const int half = warp >> 2, r = tid & (UQ - 1), i = q0 + r;
    const int cbeg = half * NC, obeg = half * NO;
    float* rl = reinterpret_cast<float*>(smem + U_OFF_RL);
    float* xm = reinterpret_cast<float*>(smem + U_OFF_XM);
    const int pair_bar = 2 + (warp & 3);      // named barrier of the two warps that share this lane quarter (NH = 2)
    // rel-key logits q_i . Ek[rel] in fp32 when this CTA's keys touch the band of its queries
    const bool any_band = kb_begin * UK < q0 + UQ + window && kb_end * UK > q0 - window;
    if (any_band) {
      float* ek = reinterpret_cast<float*>(smem + U_OFF_P);
      for (int t = tid; t < R * AD; t += U_SOFT) ek[t] = rel_k[t];
      float q[AD];
      if (i < T) {
        const float4* row = reinterpret_cast<const float4*>(qkv + ((size_t)b * T + i) * 3 * H + h * AD);
#pragma unroll
        for (int c4 = 0; c4 < AD / 4; ++c4) {
          const float4 v = row[c4];
          q[4 * c4] = v.x * qscale; q[4 * c4 + 1] = v.y * qscale; q[4 * c4 + 2] = v.z * qscale; q[4 * c4 + 3] = v.w * qscale;
        }
      } else {
#pragma unroll
        for (int c = 0; c < AD; ++c) q[c] = 0.f;
      }
      named_bar_sync(1, U_SOFT);
      for (int rel = half; rel < R; rel += NH) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;     // four chains: a single 96-term chain is latency-bound
        const float* e = ek + rel * AD;
#pragma unroll
        for (int c = 0; c < AD; c += 4) {
          s0 = fmaf(q[c], e[c], s0);
          s1 = fmaf(q[c + 1], e[c + 1], s1);
          s2 = fmaf(q[c + 2], e[c + 2], s2);
          s3 = fmaf(q[c + 3], e[c + 3], s3);
        }
        rl[r * URL + rel] = (s0 + s1) + (s2 + s3);
      }
      named_bar_sync(1, U_SOFT);     // the table's shared memory becomes the P buffer
    }
    float* band_row = band_s + (bh * Tp + i) * RP;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint8_t* p_row = smem + U_OFF_P + r * 16;
    float o[NO];
#pragma unroll
    for (int c = 0; c < NO; ++c) o[c] = 0.f;
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;

    auto fold_o = [&](int jj, float alpha) {     // o = o * alpha + O(jj)   (this thread's NO columns)
      const uint32_t t0 = lane_addr + 128u + (uint32_t)(jj & 1) * 128u + (uint32_t)obeg;
#pragma unroll
      for (int c0 = 0; c0 + 32 <= NO; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(t0 + c0, v);
#pragma unroll
        for (int c = 0; c < 32; ++c) o[c0 + c] = fmaf(o[c0 + c], alpha, __uint_as_float(v[c]));
      }
      if constexpr (NO % 32 == 16) {
        uint32_t v[16];
        tmem_ld16(t0 + (NO - 16), v);
#pragma unroll
        for (int c = 0; c < 16; ++c) o[NO - 16 + c] = fmaf(o[NO - 16 + c], alpha, __uint_as_float(v[c]));
      }
    };

    for (int jj = 0; jj < n; ++jj) {
      const int k0 = (kb_begin + jj) * UK;
      mbar_wait(&s_full[jj & 1], (jj >> 1) & 1);
      tcgen05_fence_after();
      float s[NC];
      {
        const uint32_t t0 = lane_addr + (uint32_t)(jj & 1) * UK + (uint32_t)cbeg;
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(t0 + c0, v);
#pragma unroll
          for (int c = 0; c < 32; ++c) s[c0 + c] = __uint_as_float(v[c]);
        }
      }
      const bool band = (k0 < q0 + UQ + window) && (k0 + UK > q0 - window);
      const bool edge = (k0 + UK > len) || (q0 + UQ > len) || (k0 + UK > T);   // CTA-uniform: some mask applies
      if (band || edge) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int j = k0 + cbeg + c;
          float v = s[c];
          const int rel = j - i + window;
          const bool inband = band && rel >= 0 && rel < R;
          if (inband) v += rl[r * URL + rel];
          if (!(i < len && j < len)) v = -1e4f;
          if (j >= T) v = -INFINITY;
          if (inband && j < T) band_row[rel] = v;      // each (i, j) of the band has one owner
          s[c] = v;
        }
      }
      float mx4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
      for (int c = 4; c < NC; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], s[c]);
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      if constexpr (NH == 2) {        // the row maximum over both column halves (double-buffered by block parity)
        float* slot = xm + (jj & 1) * (2 * UQ);
        slot[half * UQ + r] = mx;
        named_bar_sync(pair_bar, 64);
        mx = fmaxf(mx, slot[(half ^ 1) * UQ + r]);
      }
      const float mn = fmaxf(m, mx);
      const float alpha = __expf(m - mn);
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
      uint4 hi[NC / 8], lo[NC / 8];
#pragma unroll
      for (int g = 0; g < NC / 8; ++g) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p0 = __expf(s[g * 8 + 2 * e] - mn);       // same form as the alpha / merge weights
          const float p1 = __expf(s[g * 8 + 2 * e + 1] - mn);
          rs4[e] += p0 + p1;
          const __half2 h2 = __floats2half2_rn(p0, p1);
          const float2 hf = __half22float2(h2);
          const __half2 l2 = __floats2half2_rn(p0 - hf.x, p1 - hf.y);
          hw[e] = *reinterpret_cast<const uint32_t*>(&h2);
          lw[e] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        hi[g] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        lo[g] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
      l = fmaf(l, alpha, (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));      // NH = 2: this thread's half of the row sum
      m = mn;
      if (jj > 0) {                     // PV(jj-1) has finished reading the P buffer (and O(jj-1) is complete)
        mbar_wait(&o_full[(jj - 1) & 1], ((jj - 1) >> 1) & 1);
        tcgen05_fence_after();
      }
#pragma unroll
      for (int g = 0; g < NC / 8; ++g) {
        const int kg = cbeg / 8 + g;
        *reinterpret_cast<uint4*>(p_row + kg * (UQ * 16)) = hi[g];
        *reinterpret_cast<uint4*>(p_row + UP_TERM + kg * (UQ * 16)) = lo[g];
      }
      fence_proxy_async_smem();         // generic-proxy stores -> visible to the tensor core's async-proxy reads
      tcgen05_fence_before();           // orders this thread's tcgen05.ld of S(jj) / O(jj-2) before the arrive
      mbar_arrive(p_full);
      if (jj > 0) fold_o(jj - 1, alpha_prev);
      alpha_prev = alpha;
    }
    mbar_wait(&o_full[(n - 1) & 1], ((n - 1) >> 1) & 1);
    tcgen05_fence_after();
    fold_o(n - 1, alpha_prev);
    if constexpr (NH == 2) {            // the two halves' row sums meet once
      float* slot = xm + (n & 1) * (2 * UQ);
      slot[half * UQ + r] = l;
      named_bar_sync(pair_bar, 64);
      l += slot[(half ^ 1) * UQ + r];
    }
    // this split's running max / sum and unnormalised output
    const size_t row = (size_t)sp * rows_all + bh * Tp + i;
    if (half == 0) {
      part_m[row] = m;
      part_l[row] = l;
    }
    float4* dst = reinterpret_cast<float4*>(part_o + row * AD + obeg);
#pragma unroll
    for (int c4 = 0; c4 < NO / 4; ++c4) dst[c4] = make_float4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == MMA_W) tcgen05_dealloc(tmem, U_TMEM_COLS);
}

// split count of the tcgen05 kernel: 128-query tiles, one CTA per SM, a CTA's time ~ (its key blocks + 2)
int attn_splits_umma(int T, int heads) {
  const int nblk = (T + UQ - 1) / UQ * (UQ / UK);
  const int tiles = (T + UQ - 1) / UQ * heads, slots = device_sm_count();
  int best = 1, best_cost = 1 << 30;
  for (int ns = 1; ns <= 16 && ns <= nblk; ++ns) {
    const int waves = (tiles * ns + slots - 1) / slots;
    const int cost = waves * ((nblk + ns - 1) / ns + 2);
    if (cost < best_cost) {
      best_cost = cost;
      best = ns;
    }
  }
  return best;
}

// Key-range splits per query tile.  The grid is tiles*splits CTAs on 2 resident slots per SM and a CTA's
// time is ~(its key blocks + 1 for prologue/epilogue): pick the split count with the fewest
// wave-steps, e.g. T = 3435 (108 tiles, 54 key blocks): 5 splits = 2 waves x 12 instead of 3 splits =
// 2 waves x 19.  Depends on T only (not on B), so a row's result is independent of its batch.
int attn_splits(int T, int heads) {
  const int nblk = (T + AK - 1) / AK;
  const int tiles = nblk * heads, slots = 2 * device_sm_count();
  int best = 1, best_cost = 1 << 30;
  for (int ns = 1; ns <= 16 && ns <= nblk; ++ns) {
    const int waves = (tiles * ns + slots - 1) / slots;
    const int cost = waves * ((nblk + ns - 1) / ns + 1);
    if (cost < best_cost) {
      best_cost = cost;
      best = ns;
    }
  }
  return best;
}

}  // namespace

size_t rel_attention_scratch_bytes(int B, int T, int H) {
  const int heads = H / AD > 0 ? H / AD : 1;
  // q/k/v hi+lo (f16), then per split: partial output, max, sum (f32), then the band scores (f32); sized for
  // whichever kernel needs more (tcgen05: Tp rounded to 128; mma.sync twin: to 64, possibly more splits)
  size_t best = 0;
  for (int legacy = 0; legacy < 2; ++legacy) {
    const int q = legacy ? AK : UQ;
    const int Tp = (T + q - 1) / q * q;
    const size_t rows = (size_t)B * heads * Tp;
    const size_t ns = (size_t)(legacy ? attn_splits(T, heads) : attn_splits_umma(T, heads));
    const size_t bytes = (size_t)6 * B * H * Tp * sizeof(__half) + (ns * rows * (AD + 2) + rows * RP) * sizeof(float) + 512;
    best = bytes > best ? bytes : best;
  }
  return best;
}

cudaError_t launch_rel_attention_mma(const float* qkv, const float* rel_k, const float* rel_v, const int* lens,
                                     float* out, void* scratch, int B, int T, int H, int n_heads, int window,
                                     int legacy, cudaStream_t s) {
  if (H / n_heads != AD || H % n_heads || 2 * window + 1 > RP) return cudaErrorInvalidValue;
  if (!legacy) {
    const int Tp = (T + UQ - 1) / UQ * UQ;
    const size_t n = (size_t)B * H * Tp;
    __half* base = reinterpret_cast<__half*>(scratch);
    __half *qb = base, *kb = base + 2 * n, *vb = base + 4 * n;
    const int ns = attn_splits_umma(T, n_heads);
    const size_t rows = (size_t)B * n_heads * Tp;
    float* part_o = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + ((6 * n * sizeof(__half) + 255) & ~size_t(255)));
    float* part_m = part_o + (size_t)ns * rows * AD;
    float* part_l = part_m + (size_t)ns * rows;
    float* band_s = part_l + (size_t)ns * rows;
    const float qscale = rsqrtf((float)AD);
    attn_prep_umma_kernel<<<dim3(Tp / UK, n_heads, B), 256, 0, s>>>(qkv, qb, kb, vb, T, Tp, H, n_heads, qscale);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    static DeviceOnce once;
    static DeviceOnce once1;
    const char* nh_env = getenv("PG_ATTN_NH");      // A/B aid: 1 = one softmax thread per row, 2 (default) = two
    const bool two = !nh_env || atoi(nh_env) != 1;
    e = two ? ensure_dyn_smem(rel_attention_umma_kernel<2>, once, U_SMEM)
            : ensure_dyn_smem(rel_attention_umma_kernel<1>, once1, U_SMEM);
    if (e != cudaSuccess) return e;
    const dim3 grid(Tp / UQ, n_heads, B * ns);
    if (two)
      rel_attention_umma_kernel<2><<<grid, 320, U_SMEM, s>>>(qb, kb, vb, qkv, rel_k, lens, part_o, part_m, part_l, band_s, B,
                                                            T, Tp, H, n_heads, window, ns, qscale,
                                                            umma::make_idesc(UQ, UK), umma::make_idesc(UQ, AD));
    else
      rel_attention_umma_kernel<1><<<grid, 192, U_SMEM, s>>>(qb, kb, vb, qkv, rel_k, lens, part_o, part_m, part_l, band_s, B,
                                                            T, Tp, H, n_heads, window, ns, qscale,
                                                            umma::make_idesc(UQ, UK), umma::make_idesc(UQ, AD));
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t warps = (size_t)B * n_heads * T;
    attn_merge_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(part_o, part_m, part_l, band_s, rel_v, lens, out, B, T,
                                                                 Tp, H, n_heads, window, ns);
    return cudaGetLastError();
  }
  const int Tp = (T + AK - 1) / AK * AK;
  const size_t n = (size_t)B * H * Tp;
  __half* base = reinterpret_cast<__half*>(scratch);
  __half *qh = base, *ql = base + n, *kh = base + 2 * n, *kl = base + 3 * n, *vth = base + 4 * n, *vtl = base + 5 * n;
  const int ns = attn_splits(T, n_heads);
  const size_t rows = (size_t)B * n_heads * Tp;
  float* part_o = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + ((6 * n * sizeof(__half) + 255) & ~size_t(255)));
  float* part_m = part_o + (size_t)ns * rows * AD;
  float* part_l = part_m + (size_t)ns * rows;
  float* band_s = part_l + (size_t)ns * rows;
  const float qscale = rsqrtf((float)AD);
  attn_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(qkv, qh, ql, kh, kl, vth, vtl, B, T, Tp, H, n_heads,
                                                                qscale);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  static DeviceOnce once;
  e = ensure_dyn_smem(rel_attention_mma_kernel, once, (int)sizeof(AttnSmem));
  if (e != cudaSuccess) return e;
  dim3 grid(Tp / AQ, n_heads, B * ns);
  rel_attention_mma_kernel<<<grid, ATT_THREADS, sizeof(AttnSmem), s>>>(qh, ql, kh, kl, vth, vtl, qkv, rel_k, lens,
                                                                       part_o, part_m, part_l, band_s, B, T, Tp, H,
                                                                       n_heads, window, ns, qscale);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const size_t warps = (size_t)B * n_heads * T;
  attn_merge_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(part_o, part_m, part_l, band_s, rel_v, lens, out, B, T,
                                                               Tp, H, n_heads, window, ns);
  return cudaGetLastError();
}

}  // namespace pg
