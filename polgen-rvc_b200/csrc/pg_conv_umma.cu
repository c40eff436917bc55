// tcgen05 / TMEM implicit-GEMM Conv1d over time-major f16 activations (sm_100a).
//
// Computes one layer of the GeneratorNSF hot loop (reference
// rvc/lib/algorithm/residuals.py:45-53, conv factory :15-25):
//   y[b][t][co] = epi( sum_{tap,ci} W[tap][co][ci] * lrelu(x[b][t + tap*dil - pad][ci]) )
//
// Mapping: D[M = 128 time rows, N = Cout] += A * B per tap and
// per 16-channel K slice, accumulated in TMEM (fp32).
//   A = activations.  The CTA stages the window of 128 + (K-1)*dil rows ONCE in
//       shared memory as channel-chunk planes [Cin/8][rows][8 ch] (the UMMA
//       no-swizzle K-major canonical layout with SBO = 128 B), applying the
//       pre-activation leaky-ReLU on the way in.  Rows are 16 B apart inside a
//       plane, so a tap/dilation shift is just `start_address += shift*16`:
//       zero-copy implicit im2col, every tap re-reads the same window.
//   B = weights, [tap][Cout][Cin] f16 in global, streamed by TMA (128B swizzle,
//       64-channel chunks; 64B swizzle for Cin = 32) through an mbarrier ring.
// Roles: warps 0-3 stage A then run the epilogue (TMEM -> regs -> bias /
// residual / scale / accumulate / leaky-ReLU -> f16 -> global), warp 4 is the
// TMA producer, warp 5 owns TMEM and issues the MMAs (one elected lane).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <map>
#include <mutex>
#include <tuple>

#include "pg_common.cuh"

namespace pg {

namespace {

constexpr int BM = 128;              // output rows (time) per CTA
constexpr int NTHREADS = 192;
constexpr int MAX_WINDOW = BM + 10 * 5;   // K <= 11, dil <= 5

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
      printf("pg_conv_umma: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x,
             blockIdx.y, threadIdx.x);
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
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (f16/bf16 operands, fp32 accumulate)
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
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct UmmaParams {
  const __half* x; const __half* res; __half* y;
  const float* bias;
  int B, L;                 // L_in == L_out
  int K, dil, pad;
  float in_slope, out_slope, out_scale, res_scale;
  int act, accumulate;
  int rows_alloc;           // window rows rounded up to 8
  int plane_bytes;          // A plane pitch (16 B * (rows_alloc + 1))
  int stages;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(NTHREADS)
conv_umma_kernel(const __grid_constant__ CUtensorMap wmap, const UmmaParams p) {
  constexpr int KC = CIN >= 64 ? 64 : CIN;          // channels per weight stage
  constexpr int NCHUNK = CIN / KC;
  constexpr int STAGE_BYTES = COUT * KC * 2;
  constexpr uint32_t W_LAYOUT = KC == 64 ? LAYOUT_SW128 : LAYOUT_SW64;
  constexpr uint32_t W_SBO = 8 * KC * 2;
  constexpr int TMEM_COLS = COUT < 32 ? 32 : COUT;
  constexpr int PLANES = CIN / 8;
  constexpr uint32_t IDESC = make_idesc(BM, COUT);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [W ring | A planes | barriers]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_ring = smem;
  uint8_t* a_planes = w_ring + (size_t)p.stages * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_planes + (size_t)PLANES * p.plane_bytes);
  uint64_t* full = bars;                  // [stages]
  uint64_t* empty = bars + p.stages;      // [stages]
  uint64_t* a_ready = bars + 2 * p.stages;
  uint64_t* d_ready = a_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_ready + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * BM;

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(a_ready, 128);
    mbar_init(d_ready, 1);
    fence_barrier_init();
  }
  if (warp == 5) tcgen05_alloc(tmem_slot, TMEM_COLS);
  if (warp == 4 && lane == 0)
    asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_iters = NCHUNK * p.K;

  if (warp == 4) {
    // ===== TMA producer: weights, (chunk, tap) order =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_iters; ++it) {
        const int chunk = it / p.K, tap = it % p.K;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], STAGE_BYTES);
        tma_load_2d(w_ring + (size_t)stage * STAGE_BYTES, &wmap, &full[stage], chunk * KC, tap * COUT);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      mbar_wait(a_ready, 0);
      tcgen05_fence_after();
      const uint32_t a_base = smem_u32(a_planes);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_iters; ++it) {
        const int chunk = it / p.K, tap = it % p.K;
        mbar_wait(&full[stage], phase);
        tcgen05_fence_after();
        const uint32_t w_base = smem_u32(w_ring + (size_t)stage * STAGE_BYTES);
        const uint32_t row_shift = (uint32_t)(tap * p.dil) * 16u;
#pragma unroll
        for (int k16 = 0; k16 < KC / 16; ++k16) {
          const uint32_t plane = (uint32_t)(chunk * (KC / 8) + 2 * k16);
          const uint64_t adesc = make_desc(a_base + plane * p.plane_bytes + row_shift,
                                           (uint32_t)p.plane_bytes, 128u, LAYOUT_NONE);
          const uint64_t bdesc = make_desc(w_base + k16 * 32, 0u, W_SBO, W_LAYOUT);
          umma_f16(tmem_base, adesc, bdesc, IDESC, (it | k16) != 0);
        }
        tcgen05_commit(&empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      tcgen05_commit(d_ready);
    }
  } else {
    // ===== warps 0-3: stage the activation window, then epilogue =====
    {
      const int rows = BM + (p.K - 1) * p.dil;
      const int nvec = rows * PLANES;
      const __half* xb = p.x + (size_t)b * p.L * CIN;
      const float slope = p.in_slope;
      for (int v = tid; v < nvec; v += 128) {
        const int r = v / PLANES, pl = v % PLANES;
        const int t = t0 - p.pad + r;
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        if (t >= 0 && t < p.L) {
          q = *reinterpret_cast<const uint4*>(xb + (size_t)t * CIN + pl * 8);
          if (slope != 1.f) {
            __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float2 f = __half22float2(h[i]);
              f.x = f.x > 0.f ? f.x : f.x * slope;
              f.y = f.y > 0.f ? f.y : f.y * slope;
              h[i] = __floats2half2_rn(f.x, f.y);
            }
          }
        }
        *reinterpret_cast<uint4*>(a_planes + (size_t)pl * p.plane_bytes + (size_t)r * 16) = q;
      }
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(a_ready);
    }
    mbar_wait(d_ready, 0);
    tcgen05_fence_after();
    const int row = t0 + warp * 32 + lane;
    const bool row_ok = row < p.L;
    const size_t goff = ((size_t)b * p.L + (row_ok ? row : 0)) * COUT;
#pragma unroll 1
    for (int c0 = 0; c0 < COUT; c0 += 32) {
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, acc);
      if (!row_ok) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + c0 + j);
      }
      if (p.res) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res + goff + c0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 rv = rp[q];
          const __half2* h = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h[i]);
            v[q * 8 + 2 * i] += p.res_scale * f.x;
            v[q * 8 + 2 * i + 1] += p.res_scale * f.y;
          }
        }
      }
      if (p.out_scale != 1.f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
      }
      uint4* yp = reinterpret_cast<uint4*>(p.y + goff + c0);
      if (p.accumulate) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 ov = yp[q];
          const __half2* h = reinterpret_cast<const __half2*>(&ov);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h[i]);
            v[q * 8 + 2 * i] += f.x;
            v[q * 8 + 2 * i + 1] += f.y;
          }
        }
      }
      if (p.act == ACT_LRELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.out_slope;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 ov;
        __half2* h = reinterpret_cast<__half2*>(&ov);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[q * 8 + 2 * i], v[q * 8 + 2 * i + 1]);
        yp[q] = ov;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) tcgen05_dealloc(tmem_base, TMEM_COLS);
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
  const void* ptr; int cin, cout, k;
  bool operator<(const MapKey& o) const {
    return std::tie(ptr, cin, cout, k) < std::tie(o.ptr, o.cin, o.cout, o.k);
  }
};
std::map<MapKey, CUtensorMap> g_maps;
std::mutex g_maps_mu;

// weights [K][Cout][Cin] f16 viewed as a 2-D tensor {Cin, K*Cout}
bool get_wmap(const void* w16, int cin, int cout, int k, CUtensorMap* out) {
  std::lock_guard<std::mutex> lk(g_maps_mu);
  MapKey key{w16, cin, cout, k};
  auto it = g_maps.find(key);
  if (it != g_maps.end()) {
    *out = it->second;
    return true;
  }
  auto enc = get_encode();
  if (!enc) return false;
  const int kc = cin >= 64 ? 64 : cin;
  cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)k * cout};
  cuuint64_t strides[1] = {(cuuint64_t)cin * 2};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)cout};
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

template <int CIN, int COUT>
cudaError_t launch_t(const ConvArgs& a, cudaStream_t s) {
  CUtensorMap wmap;
  if (!get_wmap(a.w16, CIN, COUT, a.K, &wmap)) return cudaErrorNotSupported;
  constexpr int KC = CIN >= 64 ? 64 : CIN;
  constexpr int STAGE_BYTES = COUT * KC * 2;
  UmmaParams p;
  p.x = reinterpret_cast<const __half*>(a.x);
  p.res = reinterpret_cast<const __half*>(a.res);
  p.y = reinterpret_cast<__half*>(a.y);
  p.bias = a.bias;
  p.B = a.B; p.L = a.L_out; p.K = a.K; p.dil = a.dil; p.pad = a.pad;
  p.in_slope = a.in_slope; p.out_slope = a.out_slope; p.out_scale = a.out_scale;
  p.res_scale = a.res_scale; p.act = a.act; p.accumulate = a.accumulate;
  const int rows = BM + (a.K - 1) * a.dil;
  p.rows_alloc = (rows + 7) & ~7;
  p.plane_bytes = 16 * (p.rows_alloc + 1);
  const size_t a_bytes = (size_t)(CIN / 8) * p.plane_bytes;
  // as many weight stages as fit next to a second resident CTA, 2..6
  int stages = (int)((110 * 1024 - (long)a_bytes - 2048) / STAGE_BYTES);
  if (stages < 3) stages = (int)((225 * 1024 - (long)a_bytes - 2048) / STAGE_BYTES);
  if (stages > 6) stages = 6;
  if (stages > a.K * (CIN / KC)) stages = a.K * (CIN / KC);
  if (stages < 1) return cudaErrorInvalidValue;
  p.stages = stages;
  const size_t smem = 1024 + (size_t)stages * STAGE_BYTES + a_bytes + 256;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<CIN, COUT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  dim3 grid((a.L_out + BM - 1) / BM, a.B);
  conv_umma_kernel<CIN, COUT><<<grid, NTHREADS, smem, s>>>(wmap, p);
  return cudaGetLastError();
}

}  // namespace

bool umma_conv_supported(const ConvArgs& a) {
  if (!a.w16 || a.Cin != a.Cout) return false;
  if (a.Cin != 32 && a.Cin != 64 && a.Cin != 128 && a.Cin != 256) return false;
  if (a.x_ld != a.Cin || a.x_coff != 0 || a.y_ld != a.Cout || a.y_coff != 0) return false;
  if (a.res && (a.res_ld != a.Cout || a.res_coff != 0)) return false;
  if (a.in_mask || a.out_mask || a.bbias) return false;
  if (a.L_in != a.L_out) return false;
  if (a.act != ACT_NONE && a.act != ACT_LRELU) return false;
  if ((a.K - 1) * a.dil > MAX_WINDOW - BM || a.K < 1) return false;
  return true;
}

cudaError_t launch_conv_umma(const ConvArgs& a, cudaStream_t s) {
  if (!umma_conv_supported(a)) return cudaErrorInvalidValue;
  switch (a.Cin) {
    case 32: return launch_t<32, 32>(a, s);
    case 64: return launch_t<64, 64>(a, s);
    case 128: return launch_t<128, 128>(a, s);
    case 256: return launch_t<256, 256>(a, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace pg
