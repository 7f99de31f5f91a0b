// K3 -- tcgen05 / TMEM GEMMs for the coupling-MLP contractions of Float32 flows (sm_100a).
//
// Every Dense layer of `fnn` (reference src/flows/utils.jl:71-100) becomes
//   forward : Y[n, out]  = act(X[n, in] W^T + b)            tc_gemm_kernel, K-major operands
//   dgrad   : gX[n, in]  = gY[n, out] W   (* leakyrelu')     tc_gemm_kernel, K-major operands
//   wgrad   : gW^T[in, out] = X^T[in, n] gY[n, out]          tc_wgrad_kernel, MN-major operands (contraction over samples)
// Operands are bf16 "split planes": x = hi + lo with hi = bf16(x), lo = bf16(x - hi).  The parity
// mode NF_MMA_BF16X3 issues three kind::f16 MMAs per K step (hi*hi + hi*lo + lo*hi) with fp32
// accumulation in TMEM, which keeps the dropped term at 2^-16 relative; NF_MMA_BF16X1 issues hi*hi only.
//
// Kernel anatomy (one CTA per SM, persistent over M tiles of 128 samples):
//   warp 0      TMA producer   cp.async.bulk.tensor (SWIZZLE_128B) -> smem ring, mbarrier expect_tx
//   warp 1      MMA issuer     one elected thread issues tcgen05.mma, tcgen05.commit frees ring slots
//   warps 2..5  epilogue       tcgen05.ld 32x32b accumulators (double buffered in TMEM) -> fused
//                              bias/leakyrelu/tanh/split/mask/scatter -> global
#include "tc_gemm.hpp"
#include "kernels_coupling.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <map>
#include <tuple>

namespace nf {

namespace {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1u << 28)) __trap();   // a protocol bug must fail loudly instead of hanging the GPU
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B, version 1.
//   K-major : 128-byte rows (64 bf16 along K), 8-row atoms 1024 B apart (SBO); LBO unused.
//   MN-major: 128-byte rows (64 bf16 along M/N), K rows; 8-row groups SBO apart, 64-wide MN blocks LBO apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): bf16 x bf16 -> f32, M x N, majors.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
  const float ra = a - __bfloat162float(ha), rb = b - __bfloat162float(hb);
  __nv_bfloat162 h; h.x = ha; h.y = hb;
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = pack_bf16(ra, rb);
}

// ---------------------------------------------------------------------------------------------
// forward / dgrad GEMM:  D[128 x BN] = A[128 x K] * B[BN x K]^T  (both K-major), fused epilogues
// ---------------------------------------------------------------------------------------------
enum : int { EPI_PLANES_ACT = 0, EPI_F32_ACT = 1, EPI_PLANES_MASK = 2, EPI_SCATTER_ADD = 3 };

struct GemmParams {
  int64_t M;               // valid rows
  int num_k_chunks;        // K_pad / 64
  int terms;               // 1 or 3
  int epi;
  int n_store;             // number of output columns to store (planes: padded width; f32/scatter: valid width)
  const float* bias;       // [>= n tile] zero padded (EPI_*_ACT)
  int act;
  __nv_bfloat16* out_hi;   // planes [M_pad, out_ld]
  __nv_bfloat16* out_lo;
  int64_t out_ld;
  float* out_f32;          // [M, out_f32_ld]
  int64_t out_f32_ld;
  const __nv_bfloat16* mask_hi;   // EPI_PLANES_MASK: sign of the stashed activation
  int64_t mask_ld;
  float* G;                // EPI_SCATTER_ADD
  int ldg;
  const int* idx;
};

template <int BN> struct GemmCfg {
  static constexpr int A_PLANE = 128 * 128;                  // 128 rows x 128 B
  static constexpr int B_PLANE = BN * 128;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int STAGES = (STAGE * 4 <= 200 * 1024) ? 4 : ((STAGE * 3 <= 200 * 1024) ? 3 : 2);
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int SMEM = STAGES * STAGE + 1024 /*align*/ + 256 /*barriers*/ + BN * 4 /*bias*/;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int S = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + S * Cfg::STAGE;              // full[S], empty[S], tfull[2], tempty[2], tmem ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * S + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen_base + S * Cfg::STAGE + 8 * (2 * S + 4));
  float* s_bias = reinterpret_cast<float*>(gen_base + S * Cfg::STAGE + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t num_tiles = (p.M + 127) / 128;
  const int n0 = blockIdx.y * BN;
  const int nk = p.num_k_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  if (threadIdx.x >= 64 && p.bias) {
    for (int i = threadIdx.x - 64; i < BN; i += 128) s_bias[i] = p.bias[n0 + i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const uint32_t bytes = (p.terms > 1 ? 2u : 1u) * (Cfg::A_PLANE + Cfg::B_PLANE);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (int)(tile * 128);
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_expect_tx(full_bar(s), bytes);
          const uint32_t st = base + s * Cfg::STAGE;
          tma_load_3d(st, &tmapA, full_bar(s), kc * 64, m0, 0);
          tma_load_3d(st + 2 * Cfg::A_PLANE, &tmapB, full_bar(s), kc * 64, n0, 0);
          if (p.terms > 1) {
            tma_load_3d(st + Cfg::A_PLANE, &tmapA, full_bar(s), kc * 64, m0, 1);
            tma_load_3d(st + 2 * Cfg::A_PLANE + Cfg::B_PLANE, &tmapB, full_bar(s), kc * 64, n0, 1);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
      uint32_t it = 0, tl = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
        const uint32_t acc = tl & 1, accph = (tl >> 1) & 1;
        mbar_wait(tempty_bar(acc), accph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t st = base + s * Cfg::STAGE;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            for (int t = 0; t < p.terms; ++t) {
              const uint32_t a_addr = st + (t == 2 ? Cfg::A_PLANE : 0) + kk * 32;
              const uint32_t b_addr = st + 2 * Cfg::A_PLANE + (t == 1 ? Cfg::B_PLANE : 0) + kk * 32;
              umma_bf16(d_tmem, make_smem_desc(a_addr, 16, 1024), make_smem_desc(b_addr, 16, 1024), idesc,
                        (kc | kk | t) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else {
    // ===== epilogue warps (TMEM lane quarter = warp % 4) =====
    const int quarter = warp & 3;
    uint32_t tl = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
      const uint32_t acc = tl & 1, accph = (tl >> 1) & 1;
      mbar_wait(tfull_bar(acc), accph);
      tc_fence_after();
      const int64_t row = tile * 128 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.n_store) break;   // warp-uniform
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        if (!row_ok) continue;
        if (p.epi == EPI_PLANES_ACT || p.epi == EPI_PLANES_MASK) {
          uint32_t hi[16], lo[16];
          if (p.epi == EPI_PLANES_ACT) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float a = __uint_as_float(v[j]) + s_bias[c * 32 + j];
              float b = __uint_as_float(v[j + 1]) + s_bias[c * 32 + j + 1];
              if (p.act == ACT_LRELU) { a = a > 0.f ? a : 0.01f * a; b = b > 0.f ? b : 0.01f * b; }
              split_pair(a, b, hi[j >> 1], lo[j >> 1]);
            }
          } else {
            const uint4* mrow = reinterpret_cast<const uint4*>(p.mask_hi + row * p.mask_ld + col0);
            uint32_t mk[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const uint4 t4 = mrow[q]; mk[4 * q] = t4.x; mk[4 * q + 1] = t4.y; mk[4 * q + 2] = t4.z; mk[4 * q + 3] = t4.w; }
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const uint32_t m2 = mk[j >> 1];
              // bf16 sign bits: low half bit 15, high half bit 31; activation > 0  <=> not negative and not zero
              const bool pa = ((m2 & 0x8000u) == 0) && ((m2 & 0x7FFFu) != 0);
              const bool pb = ((m2 & 0x80000000u) == 0) && ((m2 & 0x7FFF0000u) != 0);
              const float a = __uint_as_float(v[j]) * (pa ? 1.f : 0.01f);
              const float b = __uint_as_float(v[j + 1]) * (pb ? 1.f : 0.01f);
              split_pair(a, b, hi[j >> 1], lo[j >> 1]);
            }
          }
          uint4* oh = reinterpret_cast<uint4*>(p.out_hi + row * p.out_ld + col0);
#pragma unroll
          for (int q = 0; q < 4; ++q) oh[q] = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
          if (p.terms > 1) {
            uint4* ol = reinterpret_cast<uint4*>(p.out_lo + row * p.out_ld + col0);
#pragma unroll
            for (int q = 0; q < 4; ++q) ol[q] = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
          }
        } else if (p.epi == EPI_F32_ACT) {
          float* o = p.out_f32 + row * p.out_f32_ld + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (col0 + j < p.n_store) {
              float a = __uint_as_float(v[j]) + s_bias[c * 32 + j];
              if (p.act == ACT_TANH) a = tanhf(a);
              else if (p.act == ACT_LRELU) a = a > 0.f ? a : 0.01f * a;
              o[j] = a;
            }
          }
        } else {  // EPI_SCATTER_ADD
          float* g = p.G + row * p.ldg;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.n_store) g[p.idx[col0 + j]] += __uint_as_float(v[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad GEMM: D[Kin x Nout] = X^T G, contraction over samples; both operands MN-major straight from
// the row-major split planes (no transposed copies).  One CTA reduces a strided set of 32-sample
// chunks into TMEM and flushes with double atomics into the theta-ordered accumulators.
// ---------------------------------------------------------------------------------------------
struct WgradParams {
  int64_t n;          // samples
  int kin, nout;      // valid sizes
  int mt;             // M tiles of 128 in-features
  int terms;
  double* gW;         // gsum + w_off, row-major [kin][nout]
};

template <int BN> struct WgradCfg {
  static constexpr int KS = 32;                          // samples per stage
  static constexpr int BLK = KS * 128;                   // one 64-wide MN block: 32 rows x 128 B
  static constexpr int A_PLANE = 4 * BLK;                // up to 256 in-features
  static constexpr int B_PLANE = (BN / 64) * BLK;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int STAGES = (STAGE * 4 <= 200 * 1024) ? 4 : 3;
  static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapG, WgradParams p) {
  using Cfg = WgradCfg<BN>;
  constexpr int S = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + S * Cfg::STAGE;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  const uint32_t done_bar = bars + 8u * (2 * S);
  const uint32_t tmem_slot = bars + 8u * (2 * S + 1);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen_base + S * Cfg::STAGE + 8 * (2 * S + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t num_chunks = (p.n + Cfg::KS - 1) / Cfg::KS;
  const uint32_t tmem_cols = (p.mt * BN <= 32) ? 32 : (p.mt * BN <= 64 ? 64 : (p.mt * BN <= 128 ? 128 : (p.mt * BN <= 256 ? 256 : 512)));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmapX);
    tma_prefetch_desc(&tmapG);
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const int a_blocks = p.mt * 2;   // 64-wide in-feature blocks fetched per plane

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t bytes = (p.terms > 1 ? 2u : 1u) * (a_blocks * Cfg::BLK + Cfg::B_PLANE);
      uint32_t it = 0;
      for (int64_t ch = blockIdx.x; ch < num_chunks; ch += gridDim.x, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_expect_tx(full_bar(s), bytes);
        const uint32_t st = base + s * Cfg::STAGE;
        const int r0 = (int)(ch * Cfg::KS);
        // 4-D view {64, rows, blocks, planes}: one TMA lands [block][row][64] per plane
        tma_load_4d(st, &tmapX, full_bar(s), 0, r0, 0, 0);
        tma_load_4d(st + 2 * Cfg::A_PLANE, &tmapG, full_bar(s), 0, r0, 0, 0);
        if (p.terms > 1) {
          tma_load_4d(st + Cfg::A_PLANE, &tmapX, full_bar(s), 0, r0, 0, 1);
          tma_load_4d(st + 2 * Cfg::A_PLANE + Cfg::B_PLANE, &tmapG, full_bar(s), 0, r0, 0, 1);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, BN, 1, 1);
      uint32_t it = 0;
      for (int64_t ch = blockIdx.x; ch < num_chunks; ch += gridDim.x, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t st = base + s * Cfg::STAGE;
#pragma unroll
        for (int ks = 0; ks < Cfg::KS / 16; ++ks) {
          for (int mt = 0; mt < p.mt; ++mt) {
            for (int t = 0; t < p.terms; ++t) {
              const uint32_t a_addr = st + (t == 2 ? Cfg::A_PLANE : 0) + mt * 2 * Cfg::BLK + ks * 16 * 128;
              const uint32_t b_addr = st + 2 * Cfg::A_PLANE + (t == 1 ? Cfg::B_PLANE : 0) + ks * 16 * 128;
              umma_bf16(tmem_base + mt * BN, make_smem_desc(a_addr, Cfg::BLK, 1024), make_smem_desc(b_addr, Cfg::BLK, 1024),
                        idesc, (it | ks | t) != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(done_bar);
    }
  } else {
    const bool any = (int64_t)blockIdx.x < num_chunks;
    if (any) {
      const int quarter = warp & 3;
      mbar_wait(done_bar, 0);
      tc_fence_after();
      for (int mt = 0; mt < p.mt; ++mt) {
        const int krow = mt * 128 + quarter * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + mt * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          if (c * 32 >= p.nout) break;
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          if (krow < p.kin) {
            double* g = p.gW + (int64_t)krow * p.nout + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < p.nout) atomicAdd(&g[j], (double)__uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// small helper kernels: split / gather into planes, plane column sums, weight preparation
// ---------------------------------------------------------------------------------------------
// planes [rows_pad, ld] hi then lo (lo plane at +plane_elems)
__global__ void gather_split_kernel(const float* __restrict__ X, int d, const int* __restrict__ idx, int n_idx, int64_t n,
                                    __nv_bfloat16* __restrict__ out, int ld, int64_t plane_elems) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * ld) return;
  const int64_t r = e / ld;
  const int k = (int)(e - r * ld);
  float v = 0.f;
  if (k < n_idx) v = idx ? X[r * d + idx[k]] : X[r * d + k];
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  out[e] = h;
  out[plane_elems + e] = __float2bfloat16_rn(v - __bfloat162float(h));
}

__global__ void plane_colsum_kernel(const __nv_bfloat16* __restrict__ P, int ld, int64_t plane_elems, int64_t n, int ncols,
                                    int64_t rows_per_block, int use_lo, double* __restrict__ out) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
      float v = __bfloat162float(P[r * ld + c]);
      if (use_lo) v += __bfloat162float(P[plane_elems + r * ld + c]);
      s += v;
    }
    atomicAdd(&out[c], (double)s);
  }
}

struct DensePrep {
  int64_t w_off, b_off;     // theta offsets
  int kin, nout;
  int kin_p, nf_rows;       // forward planes  Wf [nf_rows][kin_p]   : Wf[o][k] = Wt[k][o]
  int nout_p, nd_rows;      // dgrad planes    Wd [nd_rows][nout_p]  : Wd[k][o] = Wt[k][o]
  int64_t wf_off, wd_off, bias_off;   // element offsets into the bf16 pool / float bias pool
};

__global__ void prep_weights_kernel(const float* __restrict__ theta, const DensePrep* __restrict__ preps, int n_preps,
                                    __nv_bfloat16* __restrict__ pool, float* __restrict__ bias_pool) {
  const int pi = blockIdx.y;
  if (pi >= n_preps) return;
  const DensePrep d = preps[pi];
  const int64_t nf = (int64_t)d.nf_rows * d.kin_p, nd = (int64_t)d.nd_rows * d.nout_p;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nf + nd + d.nf_rows; e += (int64_t)gridDim.x * blockDim.x) {
    if (e < nf) {
      const int o = (int)(e / d.kin_p), k = (int)(e % d.kin_p);
      const float v = (o < d.nout && k < d.kin) ? theta[d.w_off + (int64_t)k * d.nout + o] : 0.f;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      pool[d.wf_off + e] = h;
      pool[d.wf_off + nf + e] = __float2bfloat16_rn(v - __bfloat162float(h));
    } else if (e < nf + nd) {
      const int64_t q = e - nf;
      const int k = (int)(q / d.nout_p), o = (int)(q % d.nout_p);
      const float v = (o < d.nout && k < d.kin) ? theta[d.w_off + (int64_t)k * d.nout + o] : 0.f;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      pool[d.wd_off + q] = h;
      pool[d.wd_off + nd + q] = __float2bfloat16_rn(v - __bfloat162float(h));
    } else {
      const int o = (int)(e - nf - nd);
      bias_pool[d.bias_off + o] = o < d.nout ? theta[d.b_off + o] : 0.f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps, per-flow state, launchers
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

struct TcState {
  std::vector<DensePrep> preps;             // flattened [layer][mlp][dense]
  std::vector<std::vector<std::vector<int>>> index;   // index[layer][mlp][dense] -> preps slot
  DensePrep* d_preps = nullptr;
  __nv_bfloat16* pool = nullptr;
  float* bias_pool = nullptr;
  int64_t pool_elems = 0, bias_elems = 0;
  std::map<std::tuple<const void*, int64_t, int64_t, int64_t, int, int>, CUtensorMap> maps;
};

inline int pick_bn(int n) { return n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : 256)); }
inline int pad64(int v) { return (int)round_up(v, 64); }

TcState* get_state(Flow& f) { return (TcState*)f.tc_state; }

int ensure_state(Flow& f) {
  if (f.tc_state) return NF_OK;
  TcState* st = new TcState();
  f.tc_state = st;
  int64_t pool = 0, bias = 0;
  st->index.resize(f.layers.size());
  for (size_t li = 0; li < f.layers.size(); ++li) {
    const LayerDesc& L = f.layers[li];
    st->index[li].resize(L.mlps.size());
    for (size_t m = 0; m < L.mlps.size(); ++m) {
      const MLPDesc& md = L.mlps[m];
      for (int i = 0; i < md.n_dense(); ++i) {
        DensePrep d{};
        d.w_off = md.w_off[i]; d.b_off = md.b_off[i];
        d.kin = md.dims[i]; d.nout = md.dims[i + 1];
        d.kin_p = pad64(d.kin); d.nout_p = pad64(d.nout);
        const bool last = (i + 1 == md.n_dense());
        const int nf_valid = last ? (int)round_up(d.nout, 16) : d.nout_p;
        d.nf_rows = nf_valid <= 256 ? pick_bn(nf_valid) : (int)round_up(nf_valid, 256);
        d.nd_rows = d.kin_p <= 256 ? pick_bn(i == 0 ? (int)round_up(d.kin, 16) : d.kin_p) : (int)round_up(d.kin_p, 256);
        d.wf_off = pool; pool += 2 * (int64_t)d.nf_rows * d.kin_p;
        d.wd_off = pool; pool += 2 * (int64_t)d.nd_rows * d.nout_p;
        d.bias_off = bias; bias += d.nf_rows;
        st->index[li][m].push_back((int)st->preps.size());
        st->preps.push_back(d);
      }
    }
  }
  st->pool_elems = pool; st->bias_elems = bias;
  NF_CUDA(cudaMalloc((void**)&st->pool, pool * sizeof(__nv_bfloat16)));
  NF_CUDA(cudaMalloc((void**)&st->bias_pool, bias * sizeof(float)));
  NF_CUDA(cudaMalloc((void**)&st->d_preps, st->preps.size() * sizeof(DensePrep)));
  NF_CUDA(cudaMemcpy(st->d_preps, st->preps.data(), st->preps.size() * sizeof(DensePrep), cudaMemcpyHostToDevice));
  return NF_OK;
}

// K-major view of split planes [rows][cols] (+ plane stride): dims {cols, rows, 2}, box {64, box_rows, 1}
int make_map_kmajor(TcState* st, const void* basep, int64_t rows, int64_t cols, int64_t plane_elems, int box_rows, CUtensorMap* out) {
  auto key = std::make_tuple(basep, rows, cols, plane_elems, box_rows, 0);
  auto itf = st->maps.find(key);
  if (itf != st->maps.end()) { *out = itf->second; return NF_OK; }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return NF_ERR_CUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
  cuuint64_t gstr[2] = {(cuuint64_t)cols * 2, (cuuint64_t)plane_elems * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(basep), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (K-major) failed: %d (rows %lld cols %lld)", (int)r, (long long)rows, (long long)cols); return NF_ERR_CUDA; }
  st->maps[key] = *out;
  return NF_OK;
}

// MN-major view for wgrad: dims {64, rows, cols/64, 2}, box {64, 32, nblocks, 1} -> smem [block][row][64]
int make_map_mnmajor(TcState* st, const void* basep, int64_t rows, int64_t cols, int64_t plane_elems, int nblocks, CUtensorMap* out) {
  auto key = std::make_tuple(basep, rows, cols, plane_elems, nblocks, 1);
  auto itf = st->maps.find(key);
  if (itf != st->maps.end()) { *out = itf->second; return NF_OK; }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return NF_ERR_CUDA; }
  cuuint64_t gdim[4] = {64, (cuuint64_t)rows, (cuuint64_t)(cols / 64), 2};
  cuuint64_t gstr[3] = {(cuuint64_t)cols * 2, 128, (cuuint64_t)plane_elems * 2};
  cuuint32_t box[4] = {64, 32, (cuuint32_t)nblocks, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(basep), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (MN-major) failed: %d (rows %lld cols %lld)", (int)r, (long long)rows, (long long)cols); return NF_ERR_CUDA; }
  st->maps[key] = *out;
  return NF_OK;
}

template <int BN>
int launch_gemm_bn(Flow& f, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, int n_tiles_n) {
  using Cfg = GemmCfg<BN>;
  auto kern = tc_gemm_kernel<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  const int64_t tiles = ceil_div(p.M, 128);
  dim3 grid((unsigned)std::min<int64_t>(tiles, std::max(1, kNumSMs / n_tiles_n)), (unsigned)n_tiles_n);
  kern<<<grid, 192, Cfg::SMEM, f.stream>>>(ma, mb, p);
  NF_LAUNCH_CHECK();
  return NF_OK;
}

int launch_gemm(Flow& f, int bn, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, int n_tiles_n) {
  switch (bn) {
    case 32: return launch_gemm_bn<32>(f, ma, mb, p, n_tiles_n);
    case 64: return launch_gemm_bn<64>(f, ma, mb, p, n_tiles_n);
    case 128: return launch_gemm_bn<128>(f, ma, mb, p, n_tiles_n);
    default: return launch_gemm_bn<256>(f, ma, mb, p, n_tiles_n);
  }
}

template <int BN>
int launch_wgrad_bn(Flow& f, const CUtensorMap& mx, const CUtensorMap& mg, const WgradParams& p) {
  using Cfg = WgradCfg<BN>;
  auto kern = tc_wgrad_kernel<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  const int64_t chunks = ceil_div(p.n, Cfg::KS);
  // at least 8 chunks (256 samples) per CTA so the atomic flush is amortised
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(kNumSMs, chunks / 8));
  kern<<<grid, 192, Cfg::SMEM, f.stream>>>(mx, mg, p);
  NF_LAUNCH_CHECK();
  return NF_OK;
}

struct Planes {
  __nv_bfloat16* p;
  int64_t rows_pad;
  int ld;
  int64_t plane_elems() const { return rows_pad * ld; }
};
inline Planes planes_of(void* buf, int64_t n, int width) { return Planes{(__nv_bfloat16*)buf, round_up(n, 128), pad64(width)}; }

}  // namespace

size_t tc_act_bytes(int64_t n, int width) { return (size_t)round_up(n, 128) * pad64(width) * 4; }
size_t tc_weight_bytes(const Flow&) { return 0; }

void tc_release(Flow& f) {
  TcState* st = get_state(f);
  if (!st) return;
  cudaFree(st->pool); cudaFree(st->bias_pool); cudaFree(st->d_preps);
  delete st;
  f.tc_state = nullptr;
}

int tc_prepare_weights(Flow& f, const float* theta_dev) {
  NF_TRY(ensure_state(f));
  TcState* st = get_state(f);
  if (st->preps.empty()) return NF_OK;
  dim3 grid(64, (unsigned)st->preps.size());
  prep_weights_kernel<<<grid, 256, 0, f.stream>>>(theta_dev, st->d_preps, (int)st->preps.size(), st->pool, st->bias_pool);
  NF_LAUNCH_CHECK();
  return NF_OK;
}

int tc_gather_split(Flow& f, const float* X, int d, const int* d_idx, int n_idx, int64_t n, void* act0) {
  Planes P = planes_of(act0, n, n_idx);
  gather_split_kernel<<<(unsigned)ceil_div(n * P.ld, 256), 256, 0, f.stream>>>(X, d, d_idx, n_idx, n, P.p, P.ld, P.plane_elems());
  NF_LAUNCH_CHECK();
  return NF_OK;
}

int tc_mlp_forward(Flow& f, const LayerDesc& Ld, int m, int64_t n, void* act0, std::vector<void*>& acts) {
  TcState* st = get_state(f);
  const int li = (int)(&Ld - f.layers.data());
  const MLPDesc& md = Ld.mlps[m];
  const int terms = f.mma_mode == NF_MMA_BF16X1 ? 1 : 3;
  const int nd = md.n_dense();
  for (int i = 0; i < nd; ++i) {
    const DensePrep& dp = st->preps[st->index[li][m][i]];
    const bool last = (i + 1 == nd);
    Planes A = planes_of(i == 0 ? act0 : acts[i - 1], n, dp.kin);
    const int bn = std::min(dp.nf_rows, 256);
    const int n_tiles_n = dp.nf_rows / bn;
    CUtensorMap ma, mb;
    NF_TRY(make_map_kmajor(st, A.p, n, A.ld, A.plane_elems(), 128, &ma));
    NF_TRY(make_map_kmajor(st, st->pool + dp.wf_off, dp.nf_rows, dp.kin_p, (int64_t)dp.nf_rows * dp.kin_p, bn, &mb));
    GemmParams p{};
    p.M = n; p.num_k_chunks = dp.kin_p / 64; p.terms = terms;
    p.bias = st->bias_pool + dp.bias_off;
    if (!last) {
      Planes O = planes_of(acts[i], n, dp.nout);
      p.epi = EPI_PLANES_ACT; p.act = ACT_LRELU; p.n_store = O.ld;
      p.out_hi = O.p; p.out_lo = O.p + O.plane_elems(); p.out_ld = O.ld;
    } else {
      p.epi = EPI_F32_ACT; p.act = md.out_act ? ACT_TANH : ACT_NONE; p.n_store = dp.nout;
      p.out_f32 = (float*)acts[i]; p.out_f32_ld = dp.nout;
    }
    NF_TRY(launch_gemm(f, bn, ma, mb, p, n_tiles_n));
  }
  return NF_OK;
}

int tc_mlp_backward(Flow& f, const LayerDesc& Ld, int m, int64_t n, void* act0, std::vector<void*>& acts, float* g_last,
                    void* scratch0, void* scratch1, float* G, double* gsum) {
  TcState* st = get_state(f);
  const int li = (int)(&Ld - f.layers.data());
  const MLPDesc& md = Ld.mlps[m];
  const int terms = f.mma_mode == NF_MMA_BF16X1 ? 1 : 3;
  const int nd = md.n_dense();
  // gradient w.r.t. the last pre-activation -> split planes
  void* gbuf = scratch0;
  void* gnext = scratch1;
  {
    Planes Gp = planes_of(gbuf, n, md.dims[nd]);
    gather_split_kernel<<<(unsigned)ceil_div(n * Gp.ld, 256), 256, 0, f.stream>>>(g_last, md.dims[nd], nullptr, md.dims[nd], n, Gp.p,
                                                                                  Gp.ld, Gp.plane_elems());
    NF_LAUNCH_CHECK();
  }
  for (int i = nd - 1; i >= 0; --i) {
    const DensePrep& dp = st->preps[st->index[li][m][i]];
    Planes Gp = planes_of(gbuf, n, dp.nout);
    Planes X = planes_of(i == 0 ? act0 : acts[i - 1], n, dp.kin);
    {  // bias gradient
      const int64_t rpb = 1024;
      plane_colsum_kernel<<<(unsigned)ceil_div(n, rpb), 256, 0, f.stream>>>(Gp.p, Gp.ld, Gp.plane_elems(), n, dp.nout, rpb,
                                                                          terms > 1 ? 1 : 0, gsum + dp.b_off);
      NF_LAUNCH_CHECK();
    }
    {  // weight gradient
      NF_REQUIRE(X.ld <= 256 && Gp.ld <= 256, "tcgen05 wgrad supports layer widths up to 256 (got %d x %d)", dp.kin, dp.nout);
      CUtensorMap mx, mg;
      const int mt = (int)ceil_div(X.ld, 128);
      NF_TRY(make_map_mnmajor(st, X.p, n, X.ld, X.plane_elems(), mt * 2, &mx));
      NF_TRY(make_map_mnmajor(st, Gp.p, n, Gp.ld, Gp.plane_elems(), Gp.ld / 64, &mg));
      WgradParams wp{};
      wp.n = n; wp.kin = dp.kin; wp.nout = dp.nout; wp.mt = mt; wp.terms = terms; wp.gW = gsum + dp.w_off;
      switch (Gp.ld) {
        case 64: NF_TRY(launch_wgrad_bn<64>(f, mx, mg, wp)); break;
        case 128: NF_TRY(launch_wgrad_bn<128>(f, mx, mg, wp)); break;
        case 192: NF_TRY(launch_wgrad_bn<192>(f, mx, mg, wp)); break;
        default: NF_TRY(launch_wgrad_bn<256>(f, mx, mg, wp)); break;
      }
    }
    {  // data gradient
      const int bn = std::min(dp.nd_rows, 256);
      const int n_tiles_n = dp.nd_rows / bn;
      CUtensorMap ma, mb;
      NF_TRY(make_map_kmajor(st, Gp.p, n, Gp.ld, Gp.plane_elems(), 128, &ma));
      NF_TRY(make_map_kmajor(st, st->pool + dp.wd_off, dp.nd_rows, dp.nout_p, (int64_t)dp.nd_rows * dp.nout_p, bn, &mb));
      GemmParams p{};
      p.M = n; p.num_k_chunks = dp.nout_p / 64; p.terms = terms;
      if (i > 0) {
        Planes O = planes_of(gnext, n, dp.kin);
        p.epi = EPI_PLANES_MASK; p.n_store = O.ld;
        p.out_hi = O.p; p.out_lo = O.p + O.plane_elems(); p.out_ld = O.ld;
        p.mask_hi = X.p; p.mask_ld = X.ld;
      } else {
        p.epi = EPI_SCATTER_ADD; p.n_store = dp.kin; p.G = G; p.ldg = f.dim; p.idx = Ld.d_idx2;
      }
      NF_TRY(launch_gemm(f, bn, ma, mb, p, n_tiles_n));
      std::swap(gbuf, gnext);
    }
  }
  return NF_OK;
}

}  // namespace nf
