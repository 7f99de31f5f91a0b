// K3 -- tcgen05 / TMEM GEMMs for the coupling-MLP contractions of Float32 flows (sm_100a).
//
// Every Dense layer of `fnn` (reference src/flows/utils.jl:71-100) becomes
//   forward : Y[n, out]  = act(X[n, in] W^T + b)            tc_gemm_kernel, K-major operands
//   dgrad   : gX[n, in]  = gY[n, out] W   (* leakyrelu')     tc_gemm_kernel, K-major operands
//   wgrad   : gW^T[in, out] = X^T[in, n] gY[n, out]          tc_wgrad_kernel, MN-major operands (contraction over samples)
// Operands are fp16 "split planes": s*x = hi + lo with hi = fp16(s*x), lo = fp16(s*x - hi), where s is an
// exact power of two chosen per tensor so that |s*x| <= 2^14 (no fp16 overflow; elements down to 2^-17 of
// the tensor maximum keep 22 significant bits).  s comes from a rigorous bound known before the producing
// kernel runs: bound(out) = amax(in) * max-row-L1(W) + max|b|, with amax(in) the EXACT maximum recorded
// by the kernel that produced `in` (atomicMax), so slack never compounds across layers.  The parity mode
// NF_MMA_F16X3 issues three kind::f16 MMAs per K step (hi*hi + hi*lo + lo*hi) with fp32 accumulation in
// TMEM (dropped term 2^-22 relative); NF_MMA_F16X1 issues hi*hi only (11-bit operands, not parity grade).
//
// Kernel anatomy (one CTA per SM, persistent over M tiles of 128 samples):
//   warp 0      TMA producer   cp.async.bulk.tensor (SWIZZLE_128B) -> smem ring, mbarrier expect_tx
//   warp 1      MMA issuer     one elected thread issues tcgen05.mma, tcgen05.commit frees ring slots
//   warps 2..5  epilogue       tcgen05.ld 32x32b accumulators (double buffered in TMEM) -> fused
//                              bias/leakyrelu/tanh/split/mask/scatter -> global
#include "tc_gemm.hpp"
#include "kernels_coupling.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstring>
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
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint expires) instead of
// spinning -- an ncu source view of the fused coupling kernel showed > 50 % of all issued instructions in these loops
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    if (done) break;
    if (++spins > (1u << 24)) __trap();   // a protocol bug must fail loudly instead of hanging the GPU
  }
}
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
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
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
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
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
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
//   K-major : 128-byte rows (64 fp16 along K), 8-row atoms 1024 B apart (SBO); LBO unused.
//   MN-major: 128-byte rows (64 fp16 along M/N), K rows; 8-row groups SBO apart, 64-wide MN blocks LBO apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): f16 x f16 -> f32, M x N, majors.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) /*D=f32*/ | (0u << 7) /*A=f16*/ | (0u << 10) /*B=f16*/ | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// power-of-two scale s with bound * s <= 2^14 (1 for degenerate bounds); exact arithmetic
__host__ __device__ __forceinline__ float pow2_scale(float bound) {
  if (!(bound > 0.f) || !(bound < 3.0e38f)) return 1.f;
  const int ex = ilogbf(bound) + 1;      // 2^ex > bound
  int sh = 14 - ex;
  sh = sh > 100 ? 100 : (sh < -100 ? -100 : sh);
  return ldexpf(1.f, sh);
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// per-tensor metadata: meta[0] = scale s (stored = true * s), meta[1] = bits of the exact max |true|
__device__ __forceinline__ void meta_amax(float* meta, float v) {
  atomicMax(reinterpret_cast<unsigned int*>(meta + 1), __float_as_uint(v));   // v >= 0: uint order == float order
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
  const float* a_meta;     // scale / amax of the A planes
  const float* w_sc;       // per-Dense scalars: [0] weight scale, [1] max row L1, [2] max col L1, [3] max |b|
  int bound_dgrad;         // 0: out bound = amax*w_sc[1] + w_sc[3] ; 1: amax*w_sc[2]
  float rz_comp;           // expected relative shortfall of the RZ-accumulating tensor core for this chain length
  int slab_stages;         // K stages (of 64) accumulated in one TMEM buffer before the epilogue drains it (1 .. num_k_chunks)
  float* out_meta;         // planes outputs: scale is written, amax accumulated
  const float* bias;       // [>= n tile] zero padded (EPI_*_ACT)
  int act;
  __half* out_hi;          // planes [M_pad, out_ld]
  __half* out_lo;
  int64_t out_ld;
  float* out_f32;          // [M, out_f32_ld]
  int64_t out_f32_ld;
  const uint32_t* mask_bits;   // EPI_PLANES_MASK: packed (activation > 0) bits of the stashed activation, [M_pad, mask_ld]
  int64_t mask_ld;             // words per row
  uint32_t* out_bits;          // EPI_PLANES_ACT: packed sign bits written next to the planes
  int64_t out_bits_ld;
  int f32_tma;             // EPI_F32_ACT: leave through a TMA store of the staged 32 x 16 fp32 sub-tiles (tmapO is then the fp32 map)
  float* G;                // EPI_SCATTER_ADD
  int ldg;
  const int* idx;
  double* colsum_out;      // EPI_PLANES_MASK: column sums of the written values (= bias gradient of the layer below)
  int colsum_n;
  long long* dbg;          // optional timeline of CTA 0 (clock64): [role 0..2][event idx]
  int dbg_flags;           // experiments: 1 skip plane stores, 2 skip sign-bit stores, 4 skip split/activation math
};
#define NF_DBG(role, idx, cond) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && (cond) && (idx) < 512) p.dbg[(role) * 512 + (idx)] = clock64(); } while (0)

template <int BN> struct GemmCfg {
  static constexpr int A_PLANE = 128 * 128;                  // 128 rows x 128 B
  static constexpr int B_PLANE = BN * 128;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  // narrow tiles (BN <= 64): two CTAs per SM with two stages each instead of one CTA with four -- the same bytes in flight, but
  // one CTA's epilogue overlaps the other's loads and MMAs (NFCUDA_GEMM_CTAS is a compile-time A/B switch)
#ifndef NFCUDA_GEMM_CTAS
#define NFCUDA_GEMM_CTAS 2
#endif
  static constexpr int CTAS = (BN <= 64) ? NFCUDA_GEMM_CTAS : 1;
  static constexpr int STAGES = CTAS == 2 ? 2 : ((STAGE * 4 <= 200 * 1024) ? 4 : ((STAGE * 3 <= 200 * 1024) ? 3 : 2));
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  // epilogue: 4 warps (one per TMEM lane quarter) per column group; wide tiles use four groups = 16 warps so that
  // the per-element epilogue math has 4 warps per scheduler to hide its latencies
  static constexpr int GROUPS = BN >= 128 ? 4 : BN / 32;
  static constexpr int CPT = BN / GROUPS;                    // accumulator columns held per epilogue thread (32 or 64)
  static constexpr int THREADS = 64 + 128 * GROUPS;
  // per-warp epilogue staging tile: 32 rows x 64 B, XOR-swizzled 16-byte chunks (the SWIZZLE_64B pattern): fp16 plane
  // chunks leave by TMA store, fp32 outputs use the same bytes as 32 rows x 16 floats
  static constexpr int STG_WARP = 2048;
  static constexpr int STG_OFF = ((256 + 2 * BN * 4 + 511) / 512) * 512;   // after barriers, bias, column sums; 512 B aligned
  static constexpr int SMEM = STAGES * STAGE + STG_OFF + 4 * GROUPS * STG_WARP;   // base must be 1024 B aligned (checked)
};

// The tensor core rounds its fp32 accumulator toward zero after every MMA, so a long accumulation chain
// is biased low (tests/debug_gemm.py: -8e-7 relative for K = 256).  The chain is therefore cut at every
// K slab of 64: each slab accumulates into a fresh TMEM buffer (correction products first, while the
// accumulator is still tiny), the epilogue warps drain it and carry the running sum in registers with
// round-to-nearest adds, the same split Ootomo & Yokota use for fp32 emulation on tensor cores.
template <int BN, bool SINGLE>      // SINGLE: the whole contraction accumulates in one TMEM buffer (one slab per tile)
__global__ void __launch_bounds__(GemmCfg<BN>::THREADS, GemmCfg<BN>::CTAS)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
               const __grid_constant__ CUtensorMap tmapO, GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int S = Cfg::STAGES;
  constexpr int CPT = Cfg::CPT;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();                               // SWIZZLE_128B tiles need a 1024 B aligned window
  uint8_t* gen_base = smem_raw;
  const uint32_t bars = base + S * Cfg::STAGE;              // full[S], empty[S], tfull[2], tempty[2], tmem ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * S + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen_base + S * Cfg::STAGE + 8 * (2 * S + 4));
  float* s_bias = reinterpret_cast<float*>(gen_base + S * Cfg::STAGE + 256);
  float* s_colsum = s_bias + BN;
  uint8_t* s_stage = gen_base + S * Cfg::STAGE + Cfg::STG_OFF;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t num_tiles = (p.M + 127) / 128;
  const int n0 = blockIdx.y * BN;
  const int nk = p.num_k_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4 * Cfg::GROUPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  if (threadIdx.x >= 64) {
    for (int i = threadIdx.x - 64; i < BN; i += Cfg::THREADS - 64) {
      s_bias[i] = p.bias ? p.bias[n0 + i] : 0.f;
      s_colsum[i] = 0.f;
    }
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
          NF_DBG(0, 2 * it, true);
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
    // ===== MMA issuer: one K slab (= one smem stage) per TMEM buffer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
      uint32_t it = 0;   // stage counter
      uint32_t sl = 0;   // slab counter: one TMEM buffer per slab of `slab_stages` K stages
      const int ss = p.slab_stages;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          const uint32_t acc = sl & 1, accph = (sl >> 1) & 1;
          const bool slab_first = (kc % ss) == 0;
          const bool slab_last = (kc % ss) == ss - 1 || kc == nk - 1;
          if (slab_first) mbar_wait(tempty_bar(acc), accph ^ 1);
          NF_DBG(1, 3 * it, true);
          mbar_wait(full_bar(s), ph);
          NF_DBG(1, 3 * it + 1, true);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          const uint32_t st = base + s * Cfg::STAGE;
          // one descriptor per operand per stage; every MMA below only adds a compile-time constant to its
          // 14-bit start-address field (a single thread issues these, so dependent address math is the cost)
          const uint64_t adesc0 = make_smem_desc(st, 16, 1024);
          const uint64_t bdesc0 = make_smem_desc(st + 2 * Cfg::A_PLANE, 16, 1024);
          // correction products (hi*lo, lo*hi) first, main products (hi*hi) last
          if (p.terms > 1) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)   // lo * hi
              umma_f16(d_tmem, adesc0 + ((Cfg::A_PLANE + kk * 32) >> 4), bdesc0 + ((kk * 32) >> 4), idesc, (kk > 0 || !slab_first) ? 1u : 0u);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)   // hi * lo
              umma_f16(d_tmem, adesc0 + ((kk * 32) >> 4), bdesc0 + ((Cfg::B_PLANE + kk * 32) >> 4), idesc, 1u);
          }
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)     // hi * hi
            umma_f16(d_tmem, adesc0 + ((kk * 32) >> 4), bdesc0 + ((kk * 32) >> 4), idesc, (p.terms > 1 || kk > 0 || !slab_first) ? 1u : 0u);
          umma_commit(empty_bar(s));
          if (slab_last) { umma_commit(tfull_bar(acc)); ++sl; }
          NF_DBG(1, 3 * it + 2, true);
        }
      }
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter = warp % 4, column group = (warp - 2) / 4 =====
    const int quarter = warp & 3;
    const int group = (warp - 2) >> 2;
    const int cbase = group * CPT;           // first accumulator column of this thread within the tile
    const float s_a = p.a_meta[0];
    const float amax_a = __uint_as_float(reinterpret_cast<const unsigned int*>(p.a_meta)[1]);
    const float descale0 = 1.f / (s_a * p.w_sc[0]);
    const float descale = fmaf(descale0, p.rz_comp, descale0);
    float s_out = 1.f;
    if (p.out_meta) {
      const float bound = p.bound_dgrad ? amax_a * p.w_sc[2] : amax_a * p.w_sc[1] + p.w_sc[3];
      s_out = pow2_scale(bound * 1.001f);
      if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64) p.out_meta[0] = s_out;
    }
    const float ds_out = descale * s_out;    // accumulator -> stored (scaled) value in one FFMA
    float run_max = 0.f;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      float racc[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) racc[j] = 0.f;
      const int64_t row = tile * 128 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const int ncols_here = p.n_store - (n0 + cbase);   // columns of this thread's range that matter (warp-uniform)
      // masked dgrad: the stashed sign bits of this thread's row do not depend on the accumulator -- fetch them before waiting
      // for it (one exposed global-load latency per 32-column chunk otherwise)
      uint32_t mask_pre[CPT / 32];
#pragma unroll
      for (int c = 0; c < CPT / 32; ++c) {
        const int col0 = n0 + cbase + c * 32;
        mask_pre[c] = (p.epi == EPI_PLANES_MASK && row_ok && col0 < p.n_store) ? __ldg(p.mask_bits + row * p.mask_ld + (col0 >> 5)) : 0u;
      }
      // one 32-column chunk of this thread's accumulator range (racc[c * 32 ..]): fused epilogue and store
      auto process_chunk = [&](const int c) {
        const int col0 = n0 + cbase + c * 32;
        if (col0 >= p.n_store) return;   // warp-uniform
        const float* v = racc + c * 32;
        const int bo = cbase + c * 32;     // offset into s_bias
        if (p.epi == EPI_PLANES_ACT || p.epi == EPI_PLANES_MASK) {
          uint32_t hi[16], lo[16];
          if (p.epi == EPI_PLANES_ACT) {
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              // scaled pre-activation (positive power-of-two scale commutes with leakyrelu and with the sign test)
              float a = fmaf(v[j], ds_out, s_bias[bo + j] * s_out);
              float b = fmaf(v[j + 1], ds_out, s_bias[bo + j + 1] * s_out);
              bits |= (a > 0.f ? 1u : 0u) << j;
              bits |= (b > 0.f ? 1u : 0u) << (j + 1);
              if (p.act == ACT_LRELU) { a = fmaxf(a, 0.01f * a); b = fmaxf(b, 0.01f * b); }
              run_max = fmaxf(run_max, fmaxf(fabsf(a), fabsf(b)));
              split_pair(a, b, hi[j >> 1], lo[j >> 1]);
            }
            if (row_ok && p.out_bits && !(p.dbg_flags & 2)) p.out_bits[row * p.out_bits_ld + (col0 >> 5)] = bits;
          } else {
            const uint32_t mbits = mask_pre[c];
            float cs[32];
            const float dpos = row_ok ? ds_out : 0.f, dneg = 0.01f * dpos;   // rows past M contribute exact zeros
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const bool pa = (mbits >> j) & 1u;          // leakyrelu'(h): 1 where the stashed activation was > 0, else 0.01
              const bool pb = (mbits >> (j + 1)) & 1u;
              const float a = v[j] * (pa ? dpos : dneg);                       // scaled values
              const float b = v[j + 1] * (pb ? dpos : dneg);
              run_max = fmaxf(run_max, fmaxf(fabsf(a), fabsf(b)));
              split_pair(a, b, hi[j >> 1], lo[j >> 1]);
              cs[j] = a; cs[j + 1] = b;
            }
            if (p.colsum_out) {
              // column sums over the 32 rows of this warp: recursive halving, 31 shuffles; lane l ends with column l
#pragma unroll
              for (int sft = 16; sft > 0; sft >>= 1) {
                const bool up = (lane & sft) != 0;
#pragma unroll
                for (int i = 0; i < sft; ++i) {
                  const float send = up ? cs[i] : cs[i + sft];
                  const float keep = up ? cs[i + sft] : cs[i];
                  cs[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
                }
              }
              atomicAdd(&s_colsum[cbase + c * 32 + lane], cs[0] / s_out);
            }
          }
          {
            // fp16 chunk (32 rows x 64 B) -> per-warp smem tile in the SWIZZLE_64B pattern -> one TMA store per plane;
            // rows past the end of the batch are clipped by the tensor map
            uint8_t* stg = s_stage + (warp - 2) * Cfg::STG_WARP;
            const uint32_t stg_u32 = smem_u32(stg);
            const int row_base = (int)(tile * 128) + quarter * 32;
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
              if (pl == 1 && p.terms == 1) break;
              const uint32_t* src = pl == 0 ? hi : lo;
              if (lane == 0) tma_store_wait_read();      // the previous store has finished reading the tile
              __syncwarp();
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(stg + lane * 64 + ((q ^ sw) << 4)) = make_uint4(src[4 * q], src[4 * q + 1], src[4 * q + 2], src[4 * q + 3]);
              fence_proxy_async();
              __syncwarp();
              if (lane == 0 && !(p.dbg_flags & 1)) tma_store_3d(&tmapO, stg_u32, col0, row_base, pl);
            }
          }
        } else {
          // fp32 outputs: two 16-column halves through the staging tile (64 B pitch, XOR-swizzled 16 B chunks),
          // then 4 lanes cover 64 contiguous bytes of one row
          uint8_t* stg = s_stage + (warp - 2) * Cfg::STG_WARP;
          const int64_t row_base = tile * 128 + quarter * 32;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int cc0 = col0 + hf * 16;
            if (cc0 >= p.n_store) break;     // warp-uniform
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float o4[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int j = hf * 16 + q * 4 + u;
                float a = v[j] * descale;
                if (p.epi == EPI_F32_ACT) {
                  a += s_bias[bo + j];
                  if (p.act == ACT_TANH) a = tanhf(a);
                  else if (p.act == ACT_LRELU) a = a > 0.f ? a : 0.01f * a;
                }
                o4[u] = a;
              }
              *reinterpret_cast<float4*>(stg + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = make_float4(o4[0], o4[1], o4[2], o4[3]);
            }
            if (p.epi == EPI_F32_ACT && p.f32_tma) {
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) tma_store_3d(&tmapO, smem_u32(stg), cc0, (int)row_base, 0);   // clipped at M rows / n_store columns
              continue;
            }
            __syncwarp();
            const int cq = cc0 + (lane & 3) * 4;
            if (p.epi == EPI_F32_ACT) {
#pragma unroll
              for (int ps = 0; ps < 4; ++ps) {
                const int rr = ps * 8 + (lane >> 2);
                const float4 val = *reinterpret_cast<const float4*>(stg + rr * 64 + (((lane & 3) ^ ((rr >> 1) & 3)) << 4));
                const int64_t grow = row_base + rr;
                if (grow < p.M) {
                  const float vv[4] = {val.x, val.y, val.z, val.w};
                  float* o = p.out_f32 + grow * p.out_f32_ld + cq;
                  if (cq + 3 < p.n_store && (p.out_f32_ld & 3) == 0) *reinterpret_cast<float4*>(o) = val;
                  else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (cq + u < p.n_store) o[u] = vv[u];
                  }
                }
              }
            } else {  // EPI_SCATTER_ADD: G[row, idx[col]] += v.  idx is injective, so the 16 read-modify-writes of a lane are
                      // independent: all loads are issued before the first store (one HBM round trip instead of sixteen)
              int ci[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) ci[u] = (cq + u < p.n_store) ? p.idx[cq + u] : -1;
              float old[4][4], add[4][4];
#pragma unroll
              for (int ps = 0; ps < 4; ++ps) {
                const int rr = ps * 8 + (lane >> 2);
                const int64_t grow = row_base + rr;
                const float4 val = *reinterpret_cast<const float4*>(stg + rr * 64 + (((lane & 3) ^ ((rr >> 1) & 3)) << 4));
                add[ps][0] = val.x; add[ps][1] = val.y; add[ps][2] = val.z; add[ps][3] = val.w;
                const float* g = p.G + grow * p.ldg;
#pragma unroll
                for (int u = 0; u < 4; ++u) old[ps][u] = (grow < p.M && ci[u] >= 0) ? __ldcg(g + ci[u]) : 0.f;
              }
#pragma unroll
              for (int ps = 0; ps < 4; ++ps) {
                const int64_t grow = row_base + ps * 8 + (lane >> 2);
                float* g = p.G + grow * p.ldg;
#pragma unroll
                for (int u = 0; u < 4; ++u) if (grow < p.M && ci[u] >= 0) g[ci[u]] = old[ps][u] + add[ps][u];
              }
            }
            __syncwarp();
          }
        }
      };
      const int nslabs = (nk + p.slab_stages - 1) / p.slab_stages;
      if constexpr (SINGLE) {
        // the whole contraction sits in one TMEM buffer (every backward GEMM, every K <= 64 forward GEMM): each 32-column chunk
        // is loaded right before its epilogue instead of all chunks up front, which halves the registers held through the
        // per-chunk math (no spills); the buffer goes back to the MMA warp after the last load
        const uint32_t acc = it & 1, accph = (it >> 1) & 1;
        mbar_wait_relaxed(tfull_bar(acc), accph);
        NF_DBG(2, 3 * it, threadIdx.x == 64);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + cbase;
        bool released = false;
#pragma unroll
        for (int c = 0; c < CPT / 32; ++c) {
          if (c * 32 < ncols_here) {
            uint32_t v[32];
            tmem_ld32(taddr + c * 32, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) racc[c * 32 + j] = __uint_as_float(v[j]);
          }
          if (!released && (c == CPT / 32 - 1 || (c + 1) * 32 >= ncols_here)) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            NF_DBG(2, 3 * it + 1, threadIdx.x == 64);
            released = true;
          }
          process_chunk(c);
        }
        ++it;
      } else {
        for (int kc = 0; kc < nslabs; ++kc, ++it) {
          const uint32_t acc = it & 1, accph = (it >> 1) & 1;
          mbar_wait_relaxed(tfull_bar(acc), accph);
          NF_DBG(2, 3 * it, threadIdx.x == 64);
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + cbase;
  #pragma unroll
          for (int c = 0; c < CPT / 32; ++c) {
            if (c * 32 < ncols_here) {
              uint32_t v[32];
              tmem_ld32(taddr + c * 32, v);
  #pragma unroll
              for (int j = 0; j < 32; ++j) racc[c * 32 + j] += __uint_as_float(v[j]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
          NF_DBG(2, 3 * it + 1, threadIdx.x == 64);
        }
#pragma unroll
        for (int c = 0; c < CPT / 32; ++c) process_chunk(c);
      }
      NF_DBG(2, 3 * (it - 1) + 2, threadIdx.x == 64);
    }
    if (p.out_meta) {
      run_max = warp_max(run_max) / s_out;     // tracked on the scaled values
      if (lane == 0) meta_amax(p.out_meta, run_max);
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
  if (p.colsum_out && threadIdx.x >= 64) {
    for (int i = threadIdx.x - 64; i < BN; i += Cfg::THREADS - 64)
      if (n0 + i < p.colsum_n) atomicAdd(&p.colsum_out[n0 + i], (double)s_colsum[i]);
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
  const float* x_meta;
  const float* g_meta;
  double* gW;         // gsum + w_off (+ sub-block offset), row-major with row stride ldw
  int ldw;            // row stride of gW = nout of the whole Dense (kin / nout above are the sizes of this sub-block)
  double* colsum;     // optional: bias gradient = column sums of G over the samples (gsum + b_off + sub-block offset)
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
    // with colsum the epilogue warps that own a 64-column block of G read every stage too
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1 + (p.colsum ? (BN / 64 < 4 ? BN / 64 : 4) : 0)); }
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
              umma_f16(tmem_base + mt * BN, make_smem_desc(a_addr, Cfg::BLK, 1024), make_smem_desc(b_addr, Cfg::BLK, 1024),
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
    // ---- bias gradient while the MMAs run: these warps are idle until the accumulator is complete, and the G planes of every
    // stage are already in shared memory.  Warp w owns 64-column block w of G ([32 samples][128 B], SWIZZLE_128B: 16-byte chunk
    // index ^ (row & 7)).  Lane l reads 16-byte chunk l & 7 (8 columns) of rows (l >> 3) + 4 i: eight independent fp32 sums per
    // stage, double across stages, one cross-lane reduction at the very end.
    if (p.colsum && (warp - 2) < BN / 64) {
      const int cb = warp - 2;
      const uint32_t chunk = (uint32_t)lane & 7u, rbase = (uint32_t)lane >> 3;
      double acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.0;
      uint32_t it = 0;
      for (int64_t ch = blockIdx.x; ch < num_chunks; ch += gridDim.x, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        if (lane == 0) mbar_wait(full_bar(s), ph);
        __syncwarp();
        const uint32_t gb = base + s * Cfg::STAGE + 2 * Cfg::A_PLANE + cb * Cfg::BLK;
        float sum[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) sum[q] = 0.f;
        for (int pl = 0; pl < (p.terms > 1 ? 2 : 1); ++pl) {
#pragma unroll
          for (int i = 0; i < Cfg::KS / 4; ++i) {
            const uint32_t r = rbase + 4u * i;
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                         : "r"(gb + pl * Cfg::B_PLANE + r * 128u + ((chunk ^ (r & 7u)) << 4)));
            const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 fv = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
              sum[2 * q] += fv.x; sum[2 * q + 1] += fv.y;
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += (double)sum[q];
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));
      }
      if (any) {
        const double dg = 1.0 / (double)p.g_meta[0];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          double v = acc[q];
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          const int col = cb * 64 + (int)chunk * 8 + q;
          if (lane < 8 && col < p.nout) atomicAdd(&p.colsum[col], v * dg);
        }
      }
    }
    if (any) {
      const int quarter = warp & 3;
      const float descale = 1.f / (p.x_meta[0] * p.g_meta[0]);
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
            double* g = p.gW + (int64_t)krow * p.ldw + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < p.nout) atomicAdd(&g[j], (double)(__uint_as_float(v[j]) * descale));
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
// helper kernels: absmax, split / gather into planes, plane column sums, weight preparation
// ---------------------------------------------------------------------------------------------
// max |X[r, idx[k]]| (or X[r, k]) over r < n, k < n_idx -> meta amax
__global__ void absmax_kernel(const float* __restrict__ X, int d, const int* __restrict__ idx, int n_idx, int64_t n,
                              float* __restrict__ meta) {
  float m = 0.f;
  const int64_t total = n * n_idx;
  if (!idx && d == n_idx && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {   // dense: a flat array, 16-byte loads
    const int64_t nv = total >> 2;
    const float4* X4 = reinterpret_cast<const float4*>(X);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nv; e += (int64_t)gridDim.x * blockDim.x) {
      const float4 v = X4[e];
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (int64_t e = (nv << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
      m = fmaxf(m, fabsf(X[e]));
  } else {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = e / n_idx;
      const int k = (int)(e - r * n_idx);
      m = fmaxf(m, fabsf(idx ? X[r * d + idx[k]] : X[r * d + k]));
    }
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) meta_amax(meta, m);
}

// fp32 rows -> fp16 hi/lo planes [rows_pad, ld] (hi plane, then lo plane at +plane_elems), optionally gathering columns.
// A thread owns one pair of adjacent plane columns and walks rows grid-stride (column indices hoisted, float2 loads when
// there is no gather); amax_src[1] bounds max |X|, meta receives the scale and the same bound for downstream layers.
// colsum (no-gather use): per-column sums of the unscaled input over all rows = bias gradient of the Dense whose
// pre-activation gradient X is; accumulated per thread, reduced over the block's row lanes, one double atomic per column.
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ X, int d, const int* __restrict__ idx, int n_idx,
                                                         int64_t n, __half* __restrict__ out, int ld, int64_t plane_elems,
                                                         const float* __restrict__ amax_src, float* __restrict__ meta,
                                                         double* __restrict__ colsum) {
  __shared__ float red[512];
  const int half_ld = ld >> 1;
  const int tx = threadIdx.x % half_ld, ty = threadIdx.x / half_ld, RY = blockDim.x / half_ld;
  const unsigned int abits = reinterpret_cast<const unsigned int*>(amax_src)[1];
  const float s = pow2_scale(__uint_as_float(abits));
  if (blockIdx.x == 0 && threadIdx.x == 0) { meta[0] = s; reinterpret_cast<unsigned int*>(meta)[1] = abits; }
  const int k = 2 * tx;
  int c0 = -1, c1 = -1;
  if (k < n_idx) c0 = idx ? idx[k] : k;
  if (k + 1 < n_idx) c1 = idx ? idx[k + 1] : k + 1;
  const bool vec = !idx && (d & 1) == 0 && c1 >= 0;
  float a0 = 0.f, a1 = 0.f;
  const int64_t step = (int64_t)gridDim.x * RY;
#pragma unroll 4
  for (int64_t r = (int64_t)blockIdx.x * RY + ty; r < n; r += step) {
    const float* xr = X + r * d;
    float v0 = 0.f, v1 = 0.f;
    if (vec) { const float2 t = *reinterpret_cast<const float2*>(xr + k); v0 = t.x; v1 = t.y; }
    else { if (c0 >= 0) v0 = xr[c0]; if (c1 >= 0) v1 = xr[c1]; }
    a0 += v0; a1 += v1;
    uint32_t hi, lo;
    split_pair(v0 * s, v1 * s, hi, lo);
    const int64_t o = r * ld + k;
    *reinterpret_cast<uint32_t*>(out + o) = hi;
    *reinterpret_cast<uint32_t*>(out + plane_elems + o) = lo;
  }
  if (colsum) {
    red[2 * threadIdx.x] = a0; red[2 * threadIdx.x + 1] = a1;
    __syncthreads();
    if (ty == 0) {
      for (int q = 1; q < RY; ++q) { a0 += red[2 * (q * half_ld + tx)]; a1 += red[2 * (q * half_ld + tx) + 1]; }
      if (c0 >= 0) atomicAdd(&colsum[c0], (double)a0);
      if (c1 >= 0) atomicAdd(&colsum[c1], (double)a1);
    }
  }
}

// Same job for ungathered inputs whose width is a multiple of 4: a thread owns four adjacent columns (one 16-byte load,
// two 8-byte plane stores per row), which doubles the bytes in flight per thread -- the wide case (the spline
// conditioner's [N, (3K-1)c] gradient) is purely HBM-latency bound otherwise.
__global__ void __launch_bounds__(256) split_rows4_kernel(const float* __restrict__ X, int ncols, int64_t n, __half* __restrict__ out,
                                                          int ld, int64_t plane_elems, const float* __restrict__ amax_src,
                                                          float* __restrict__ meta, double* __restrict__ colsum) {
  __shared__ float red[1024];
  const int qld = ld >> 2;
  const int tx = threadIdx.x % qld, ty = threadIdx.x / qld, RY = blockDim.x / qld;
  const unsigned int abits = reinterpret_cast<const unsigned int*>(amax_src)[1];
  const float s = pow2_scale(__uint_as_float(abits));
  if (blockIdx.x == 0 && threadIdx.x == 0) { meta[0] = s; reinterpret_cast<unsigned int*>(meta)[1] = abits; }
  const int k = 4 * tx;
  const bool live = k < ncols;          // ncols % 4 == 0: a quad is all live or all padding
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const int64_t step = (int64_t)gridDim.x * RY;
#pragma unroll 4
  for (int64_t r = (int64_t)blockIdx.x * RY + ty; r < n; r += step) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) v = *reinterpret_cast<const float4*>(X + r * ncols + k);
    a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
    uint2 hi, lo;
    split_pair(v.x * s, v.y * s, hi.x, lo.x);
    split_pair(v.z * s, v.w * s, hi.y, lo.y);
    const int64_t o = r * ld + k;
    *reinterpret_cast<uint2*>(out + o) = hi;
    *reinterpret_cast<uint2*>(out + plane_elems + o) = lo;
  }
  if (colsum) {
    red[4 * threadIdx.x] = a0; red[4 * threadIdx.x + 1] = a1; red[4 * threadIdx.x + 2] = a2; red[4 * threadIdx.x + 3] = a3;
    __syncthreads();
    if (ty == 0 && live) {
      for (int q = 1; q < RY; ++q) {
        const float* p = red + 4 * (q * qld + tx);
        a0 += p[0]; a1 += p[1]; a2 += p[2]; a3 += p[3];
      }
      atomicAdd(&colsum[k], (double)a0); atomicAdd(&colsum[k + 1], (double)a1);
      atomicAdd(&colsum[k + 2], (double)a2); atomicAdd(&colsum[k + 3], (double)a3);
    }
  }
}

struct DensePrep {
  int64_t w_off, b_off;     // theta offsets
  int kin, nout;
  int kin_p, nf_rows;       // forward planes  Wf [nf_rows][kin_p]   : Wf[o][k] = Wt[k][o]
  int nout_p, nd_rows;      // dgrad planes    Wd [nd_rows][nout_p]  : Wd[k][o] = Wt[k][o]
  int64_t wf_off, wd_off, bias_off;   // element offsets into the fp16 pool / float bias pool
};

// one block per Dense: [0] weight scale, [1] max_o sum_k |W[o,k]|, [2] max_k sum_o |W[o,k]|, [3] max |b|
__global__ void prep_scalars_kernel(const float* __restrict__ theta, const DensePrep* __restrict__ preps, float* __restrict__ sc) {
  const DensePrep d = preps[blockIdx.x];
  const float* Wt = theta + d.w_off;   // [kin][nout]
  float wmax = 0.f, rowl1 = 0.f, coll1 = 0.f, bmax = 0.f;
  for (int o = threadIdx.x; o < d.nout; o += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < d.kin; ++k) { const float a = fabsf(Wt[(int64_t)k * d.nout + o]); s += a; wmax = fmaxf(wmax, a); }
    rowl1 = fmaxf(rowl1, s);
    bmax = fmaxf(bmax, fabsf(theta[d.b_off + o]));
  }
  for (int k = threadIdx.x; k < d.kin; k += blockDim.x) {
    float s = 0.f;
    for (int o = 0; o < d.nout; ++o) s += fabsf(Wt[(int64_t)k * d.nout + o]);
    coll1 = fmaxf(coll1, s);
  }
  __shared__ float red[4][32];
  wmax = warp_max(wmax); rowl1 = warp_max(rowl1); coll1 = warp_max(coll1); bmax = warp_max(bmax);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = wmax; red[1][w] = rowl1; red[2][w] = coll1; red[3][w] = bmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    for (int i = 1; i < nw; ++i) { red[0][0] = fmaxf(red[0][0], red[0][i]); red[1][0] = fmaxf(red[1][0], red[1][i]);
                                   red[2][0] = fmaxf(red[2][0], red[2][i]); red[3][0] = fmaxf(red[3][0], red[3][i]); }
    float* o = sc + 4 * blockIdx.x;
    o[0] = pow2_scale(red[0][0]); o[1] = red[1][0]; o[2] = red[2][0]; o[3] = red[3][0];
  }
}

__global__ void prep_weights_kernel(const float* __restrict__ theta, const DensePrep* __restrict__ preps, int n_preps,
                                    const float* __restrict__ sc, __half* __restrict__ pool, float* __restrict__ bias_pool) {
  const int pi = blockIdx.y;
  if (pi >= n_preps) return;
  const DensePrep d = preps[pi];
  const float ws = sc[4 * pi];
  const int64_t nf = (int64_t)d.nf_rows * d.kin_p, nd = (int64_t)d.nd_rows * d.nout_p;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nf + nd + d.nf_rows; e += (int64_t)gridDim.x * blockDim.x) {
    if (e < nf) {
      const int o = (int)(e / d.kin_p), k = (int)(e % d.kin_p);
      const float v = (o < d.nout && k < d.kin) ? theta[d.w_off + (int64_t)k * d.nout + o] * ws : 0.f;
      const __half h = __float2half_rn(v);
      pool[d.wf_off + e] = h;
      pool[d.wf_off + nf + e] = __float2half_rn(v - __half2float(h));
    } else if (e < nf + nd) {
      const int64_t q = e - nf;
      const int k = (int)(q / d.nout_p), o = (int)(q % d.nout_p);
      const float v = (o < d.nout && k < d.kin) ? theta[d.w_off + (int64_t)k * d.nout + o] * ws : 0.f;
      const __half h = __float2half_rn(v);
      pool[d.wd_off + q] = h;
      pool[d.wd_off + nd + q] = __float2half_rn(v - __half2float(h));
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

constexpr int kMetaSlots = 8192;

// Expected relative shortfall of one K slab accumulated by the RZ-rounding tensor core: n_main = average
// number of hi*hi MMAs per slab that carry data.  Calibrated on B200 with tests/debug_gemm.py
// (4 main MMAs: -9.0e-8, 2 main MMAs: -5.3e-8); about one fp32 ulp, applied to the de-scaling factor.
// slab_stages > 1: later stages of a slab add their correction products onto an already large accumulator, so all
// 12 MMAs of those stages truncate (c2 per MMA).
float rz_compensation(int k_valid, int n_stages, int slab_stages, float c2_chain = -1.f) {
  static float c0 = -1.f, c1 = -1.f, c2 = -1.f;
  if (c0 < 0.f) {
    const char* e0 = getenv("NFCUDA_RZ_C0");
    const char* e1 = getenv("NFCUDA_RZ_C1");
    const char* e2 = getenv("NFCUDA_RZ_C2");
    c0 = e0 ? (float)atof(e0) : 2.0e-8f;
    c1 = e1 ? (float)atof(e1) : 1.75e-8f;
    c2 = e2 ? (float)atof(e2) : 1.6e-8f;
  }
  const float n_main = (float)((k_valid + 15) / 16) / (float)(n_stages > 0 ? n_stages : 1);   // data-carrying hi*hi MMAs per stage
  const int ss = slab_stages < n_stages ? slab_stages : n_stages;
  return c0 + c1 * n_main + (c2_chain >= 0.f ? c2_chain : c2) * 3.f * n_main * (float)(ss - 1);
}
// K stages per TMEM accumulation chain.  Forward GEMMs feed the ELBO value (1e-5 budget): one stage per chain.
// Backward (dgrad) GEMMs only feed the gradient (1e-4 budget): the whole K in one chain, which lets the next
// tile's MMAs overlap this tile's epilogue (tile-level TMEM double buffering).
int slab_stages_for(bool backward, int n_stages) {
  const char* e = getenv(backward ? "NFCUDA_SLAB_BWD" : "NFCUDA_SLAB_FWD");
  int ss = e ? atoi(e) : (backward ? n_stages : 1);
  if (ss < 1) ss = 1;
  if (ss > n_stages) ss = n_stages;
  return ss;
}

struct TcState {
  std::vector<DensePrep> preps;             // flattened [layer][mlp][dense]
  std::vector<std::vector<std::vector<int>>> index;   // index[layer][mlp][dense] -> preps slot
  DensePrep* d_preps = nullptr;
  float* d_scalars = nullptr;               // [n_preps][4]
  __half* pool = nullptr;
  float* bias_pool = nullptr;
  float* meta_pool = nullptr;               // [kMetaSlots][2]
  int next_slot = 0;
  std::map<const void*, int> slot_of;
  int64_t pool_elems = 0, bias_elems = 0;
  std::map<std::tuple<const void*, int64_t, int64_t, int64_t, int, int>, CUtensorMap> maps;
  std::map<int, int*> pos2;                 // fused path: [dim] position of a column inside idx2 (or -1), per layer
};

inline int pick_bn(int n) { return n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : 256)); }
inline int pad64(int v) { return (int)round_up(v, 64); }
// Row stride (elements) of a pair of split planes holding `width` columns.  Tensors of at most 32 columns keep 64-byte rows:
// the 64-column TMA boxes of the GEMM / weight-gradient loads zero-fill the columns past the tensor map's extent, so the
// narrow tensors of a spline conditioner ([N, 8] input, [N, 32] hidden layers) and the 32-wide ends of the RealNVP conditioners
// move half the bytes.  (NFCUDA_NARROW_PLANES=0 restores 128-byte rows: A/B switch.)
inline int plane_ld(int width) {
  static const bool narrow = !(getenv("NFCUDA_NARROW_PLANES") && atoi(getenv("NFCUDA_NARROW_PLANES")) == 0);
  return (narrow && width <= 32) ? 32 : pad64(width);
}

TcState* get_state(Flow& f) { return (TcState*)f.tc_state; }

// a tensor is (re)produced into `buf`: give it a fresh zeroed meta slot
float* new_meta(TcState* st, const void* buf) {
  if (st->next_slot >= kMetaSlots) return nullptr;
  const int s = st->next_slot++;
  st->slot_of[buf] = s;
  return st->meta_pool + 2 * s;
}
float* meta_of(TcState* st, const void* buf) {
  auto it = st->slot_of.find(buf);
  return it == st->slot_of.end() ? nullptr : st->meta_pool + 2 * it->second;
}

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
  NF_CUDA(cudaMalloc((void**)&st->pool, std::max<int64_t>(pool, 1) * sizeof(__half)));
  NF_CUDA(cudaMalloc((void**)&st->bias_pool, std::max<int64_t>(bias, 1) * sizeof(float)));
  NF_CUDA(cudaMalloc((void**)&st->d_preps, std::max<size_t>(st->preps.size(), 1) * sizeof(DensePrep)));
  NF_CUDA(cudaMalloc((void**)&st->d_scalars, std::max<size_t>(st->preps.size(), 1) * 4 * sizeof(float)));
  NF_CUDA(cudaMalloc((void**)&st->meta_pool, kMetaSlots * 2 * sizeof(float)));
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
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(basep), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (K-major) failed: %d (rows %lld cols %lld)", (int)r, (long long)rows, (long long)cols); return NF_ERR_CUDA; }
  if (st->maps.size() > 4096) st->maps.clear();
  st->maps[key] = *out;
  return NF_OK;
}

// Store view of split planes: dims {cols, rows, 2}, box {32, 32, 1}, SWIZZLE_64B (matches the epilogue staging tile)
int make_map_store(TcState* st, const void* basep, int64_t rows, int64_t cols, int64_t plane_elems, CUtensorMap* out) {
  auto key = std::make_tuple(basep, rows, cols, plane_elems, 32, 2);
  auto itf = st->maps.find(key);
  if (itf != st->maps.end()) { *out = itf->second; return NF_OK; }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return NF_ERR_CUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
  cuuint64_t gstr[2] = {(cuuint64_t)cols * 2, (cuuint64_t)plane_elems * 2};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(basep), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (store) failed: %d (rows %lld cols %lld)", (int)r, (long long)rows, (long long)cols); return NF_ERR_CUDA; }
  if (st->maps.size() > 4096) st->maps.clear();
  st->maps[key] = *out;
  return NF_OK;
}

// Store view of split planes for the fused coupling kernel: dims {cols, rows, 2}, box {16, 32, 1}, SWIZZLE_32B
// (one warp's piece of a hidden-activation chunk: 32 rows x 32 bytes)
int make_map_store32(TcState* st, const void* basep, int64_t rows, int64_t cols, int64_t plane_elems, CUtensorMap* out) {
  auto key = std::make_tuple(basep, rows, cols, plane_elems, 16, 4);
  auto itf = st->maps.find(key);
  if (itf != st->maps.end()) { *out = itf->second; return NF_OK; }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return NF_ERR_CUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
  cuuint64_t gstr[2] = {(cuuint64_t)cols * 2, (cuuint64_t)plane_elems * 2};
  cuuint32_t box[3] = {16, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(basep), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (store32) failed: %d (rows %lld cols %lld)", (int)r, (long long)rows, (long long)cols); return NF_ERR_CUDA; }
  if (st->maps.size() > 4096) st->maps.clear();
  st->maps[key] = *out;
  return NF_OK;
}

// Same planes, box {32, 32, 1}, SWIZZLE_64B: one warp's piece of a hidden-activation chunk in the two-team kernel (32 rows x
// 64 bytes: half the row requests of the 32-byte boxes above for the same bytes)
int make_map_store64(TcState* st, const void* basep, int64_t rows, int64_t cols, int64_t plane_elems, CUtensorMap* out) {
  auto key = std::make_tuple(basep, rows, cols, plane_elems, 32, 5);
  auto itf = st->maps.find(key);
  if (itf != st->maps.end()) { *out = itf->second; return NF_OK; }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return NF_ERR_CUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
  cuuint64_t gstr[2] = {(cuuint64_t)cols * 2, (cuuint64_t)plane_elems * 2};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(basep), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (store64) failed: %d (rows %lld cols %lld)", (int)r, (long long)rows, (long long)cols); return NF_ERR_CUDA; }
  if (st->maps.size() > 4096) st->maps.clear();
  st->maps[key] = *out;
  return NF_OK;
}

// fp32 row-major output [rows, cols] with row stride ld: box {16 columns, 32 rows} = the per-warp staging sub-tile (64-byte rows)
int make_map_store_f32(TcState* st, const void* basep, int64_t rows, int64_t cols, int64_t ld, CUtensorMap* out) {
  auto key = std::make_tuple(basep, rows, cols * 65536 + ld, (int64_t)0, 16, 3);
  auto itf = st->maps.find(key);
  if (itf != st->maps.end()) { *out = itf->second; return NF_OK; }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return NF_ERR_CUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 1};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 4, (cuuint64_t)ld * 4 * (cuuint64_t)rows};
  cuuint32_t box[3] = {16, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(basep), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (fp32 store) failed: %d (rows %lld cols %lld ld %lld)", (int)r, (long long)rows, (long long)cols, (long long)ld); return NF_ERR_CUDA; }
  if (st->maps.size() > 4096) st->maps.clear();
  st->maps[key] = *out;
  return NF_OK;
}

// MN-major view for wgrad: dims {64, rows, cols/64, 2}, box {64, 32, nblocks, 1} -> smem [block][row][64]
// `cols` = row stride of the planes in elements; `avail_cols` = columns addressable from basep (basep may point at a
// 256-column sub-block of a wider plane).
int make_map_mnmajor(TcState* st, const void* basep, int64_t rows, int64_t cols, int64_t plane_elems, int nblocks, CUtensorMap* out,
                     int64_t avail_cols = 0) {
  if (avail_cols <= 0) avail_cols = cols;
  auto key = std::make_tuple(basep, rows, cols * 4096 + avail_cols, plane_elems, nblocks, 1);
  auto itf = st->maps.find(key);
  if (itf != st->maps.end()) { *out = itf->second; return NF_OK; }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return NF_ERR_CUDA; }
  // planes narrower than one 64-column block (plane_ld): the box still spans 64 columns, the tail is zero-filled
  cuuint64_t gdim[4] = {(cuuint64_t)std::min<int64_t>(64, cols), (cuuint64_t)rows, (cuuint64_t)std::max<int64_t>(1, avail_cols / 64), 2};
  cuuint64_t gstr[3] = {(cuuint64_t)cols * 2, 128, (cuuint64_t)plane_elems * 2};
  cuuint32_t box[4] = {64, 32, (cuuint32_t)nblocks, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(basep), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (MN-major) failed: %d (rows %lld cols %lld)", (int)r, (long long)rows, (long long)cols); return NF_ERR_CUDA; }
  if (st->maps.size() > 4096) st->maps.clear();
  st->maps[key] = *out;
  return NF_OK;
}

template <int BN>
int launch_gemm_bn(Flow& f, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const GemmParams& p, int n_tiles_n) {
  using Cfg = GemmCfg<BN>;
  const bool single = p.slab_stages >= p.num_k_chunks;
  auto kern = single ? tc_gemm_kernel<BN, true> : tc_gemm_kernel<BN, false>;
  static bool attr_set[64] = {};            // per device: function attributes belong to the context
  if (!attr_set[f.device & 63]) {
    NF_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    NF_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set[f.device & 63] = true;
  }
  const int64_t tiles = ceil_div(p.M, 128);
  dim3 grid((unsigned)std::min<int64_t>(tiles, std::max(1, Cfg::CTAS * kNumSMs / n_tiles_n)), (unsigned)n_tiles_n);
  char key[64];
  snprintf(key, sizeof(key), "tc_gemm_n%d_k%d_e%d", BN, p.num_k_chunks * 64, p.epi);
  // NFCUDA_DBG=<class>: dump the clock64 timeline of CTA 0 for the n-th (NFCUDA_DBG_SKIP) launch of that class
  static int dbg_seen = 0;
  const char* dbg_cls = getenv("NFCUDA_DBG");
  long long* d_dbg = nullptr;
  GemmParams pp = p;
  pp.dbg_flags = getenv("NFCUDA_DBG_FLAGS") ? atoi(getenv("NFCUDA_DBG_FLAGS")) : 0;
  if (dbg_cls && !strcmp(dbg_cls, key)) {
    const int skip = getenv("NFCUDA_DBG_SKIP") ? atoi(getenv("NFCUDA_DBG_SKIP")) : 0;
    if (dbg_seen++ == skip) {
      NF_CUDA(cudaMalloc((void**)&d_dbg, 5 * 512 * sizeof(long long)));
      NF_CUDA(cudaMemset(d_dbg, 0, 5 * 512 * sizeof(long long)));
      pp.dbg = d_dbg;
    }
  }
  if (f.prof.on) {
    key[strlen(key) - 3] = 0;      // profile classes ignore the epilogue kind
    f.prof.begin(key, f.stream);
  }
  kern<<<grid, Cfg::THREADS, Cfg::SMEM, f.stream>>>(ma, mb, mo, pp);
  f.prof.end(f.stream);
  NF_LAUNCH_CHECK();
  if (d_dbg) {
    std::vector<long long> h(5 * 512);
    NF_CUDA(cudaStreamSynchronize(f.stream));
    NF_CUDA(cudaMemcpy(h.data(), d_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(d_dbg);
    long long t0 = 0;
    for (auto v : h) if (v && (!t0 || v < t0)) t0 = v;
    fprintf(stderr, "[nfcuda dbg] %s grid %u tiles %lld nk %d : clock64 relative to first event (CTA 0)\n", dbg_cls, grid.x, (long long)tiles, p.num_k_chunks);
    const char* names[3] = {"producer (2*it: slot free, issue)", "mma (3*it: tmem free, smem full, issued)", "epilogue w2 (3*it: tmem full, drained, [tile finalised])"};
    for (int r = 0; r < 3; ++r) {
      fprintf(stderr, "  %s\n   ", names[r]);
      for (int i = 0; i < 96; ++i) fprintf(stderr, " %lld", h[r * 512 + i] ? h[r * 512 + i] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  return NF_OK;
}

// mo: store map of the output planes (EPI_PLANES_*); any valid map otherwise (unused)
int launch_gemm(Flow& f, int bn, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const GemmParams& p, int n_tiles_n) {
  switch (bn) {
    case 32: return launch_gemm_bn<32>(f, ma, mb, mo, p, n_tiles_n);
    case 64: return launch_gemm_bn<64>(f, ma, mb, mo, p, n_tiles_n);
    case 128: return launch_gemm_bn<128>(f, ma, mb, mo, p, n_tiles_n);
    default: return launch_gemm_bn<256>(f, ma, mb, mo, p, n_tiles_n);
  }
}

template <int BN>
int launch_wgrad_bn(Flow& f, const CUtensorMap& mx, const CUtensorMap& mg, const WgradParams& p) {
  using Cfg = WgradCfg<BN>;
  auto kern = tc_wgrad_kernel<BN>;
  static bool attr_set[64] = {};            // per device: function attributes belong to the context
  if (!attr_set[f.device & 63]) {
    NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set[f.device & 63] = true;
  }
  const int64_t chunks = ceil_div(p.n, Cfg::KS);
  // at least 8 chunks (256 samples) per CTA so the atomic flush is amortised
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(kNumSMs, chunks / 8));
  if (f.prof.on) {
    char key[64];
    snprintf(key, sizeof(key), "tc_wgrad_m%d_n%d", p.mt * 128, BN);
    f.prof.begin(key, f.stream);
  }
  kern<<<grid, 192, Cfg::SMEM, f.stream>>>(mx, mg, p);
  f.prof.end(f.stream);
  NF_LAUNCH_CHECK();
  return NF_OK;
}

struct Planes {
  __half* p;
  int64_t rows_pad;
  int ld;
  int64_t plane_elems() const { return rows_pad * ld; }
  uint32_t* bits() const { return reinterpret_cast<uint32_t*>(p + 2 * plane_elems()); }   // [rows_pad, ld/32] sign bits
  int bits_ld() const { return ld / 32; }
};
inline Planes planes_of(void* buf, int64_t n, int width) { return Planes{(__half*)buf, round_up(n, 128), plane_ld(width)}; }

// fp32 [n, ld_src] (optionally gathered columns) -> split planes.  The scale comes from `amax_src` (a bound on
// max |X| recorded by the producer of X) when given, else from an exact absmax pass.
int split_into_planes(Flow& f, TcState* st, const float* X, int d, const int* d_idx, int n_idx, int64_t n, void* buf,
                      const float* amax_src, double* colsum = nullptr) {
  NF_REQUIRE(!(colsum && d_idx), "split_into_planes: column sums are for ungathered inputs");
  Planes P = planes_of(buf, n, n_idx);
  float* meta = new_meta(st, buf);
  NF_REQUIRE(meta, "tcgen05 path: out of tensor metadata slots");
  if (!amax_src) {
    const int64_t total = n * n_idx;
    absmax_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256 * 8), 4 * kNumSMs), 256, 0, f.stream>>>(X, d, d_idx, n_idx, n, meta);
    NF_LAUNCH_CHECK();
    amax_src = meta;
  }
  if (!d_idx && d == n_idx && (n_idx & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {
    const int qld = P.ld / 4;
    const int ry = std::max(1, 256 / qld);
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, ry), 8 * kNumSMs);
    split_rows4_kernel<<<grid, qld * ry, 0, f.stream>>>(X, n_idx, n, P.p, P.ld, P.plane_elems(), amax_src, meta, colsum);
    NF_LAUNCH_CHECK();
    return NF_OK;
  }
  const int half_ld = P.ld / 2;
  const int ry = std::max(1, 256 / half_ld);
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, ry), 8 * kNumSMs);
  split_rows_kernel<<<grid, half_ld * ry, 0, f.stream>>>(X, d, d_idx, n_idx, n, P.p, P.ld, P.plane_elems(), amax_src, meta, colsum);
  NF_LAUNCH_CHECK();
  return NF_OK;
}

#include "fused_coupling.cuh"
#include "fused_coupling_w128.cuh"

}  // namespace

size_t tc_act_bytes(int64_t n, int width) {
  const size_t rows = (size_t)round_up(n, 128), ld = (size_t)plane_ld(width);
  return rows * ld * 4 /* hi + lo planes */ + rows * (ld / 32) * 4 /* packed sign bits */;
}
size_t tc_weight_bytes(const Flow&) { return 0; }

void tc_release(Flow& f) {
  TcState* st = get_state(f);
  if (!st) return;
  cudaFree(st->pool); cudaFree(st->bias_pool); cudaFree(st->d_preps); cudaFree(st->d_scalars); cudaFree(st->meta_pool);
  for (auto& kv : st->pos2) cudaFree(kv.second);
  delete st;
  f.tc_state = nullptr;
}

int tc_begin_chunk(Flow& f) {
  NF_TRY(ensure_state(f));
  TcState* st = get_state(f);
  st->next_slot = 0;
  st->slot_of.clear();
  NF_CUDA(cudaMemsetAsync(st->meta_pool, 0, kMetaSlots * 2 * sizeof(float), f.stream));
  return NF_OK;
}

int tc_prepare_weights(Flow& f, const float* theta_dev) {
  NF_TRY(ensure_state(f));
  TcState* st = get_state(f);
  NF_TRY(tc_begin_chunk(f));
  if (st->preps.empty()) return NF_OK;
  prep_scalars_kernel<<<(unsigned)st->preps.size(), 256, 0, f.stream>>>(theta_dev, st->d_preps, st->d_scalars);
  NF_LAUNCH_CHECK();
  dim3 grid(64, (unsigned)st->preps.size());
  prep_weights_kernel<<<grid, 256, 0, f.stream>>>(theta_dev, st->d_preps, (int)st->preps.size(), st->d_scalars, st->pool, st->bias_pool);
  NF_LAUNCH_CHECK();
  return NF_OK;
}

int tc_gather_split(Flow& f, const float* X, int d, const int* d_idx, int n_idx, int64_t n, void* act0, const float* amax_src) {
  return split_into_planes(f, get_state(f), X, d, d_idx, n_idx, n, act0, amax_src);
}

int tc_planes_out(Flow& f, void* buf, int64_t n, int width, TcPlanesOut* out) {
  TcState* st = get_state(f);
  NF_REQUIRE(st, "tcgen05 path: no state");
  Planes P = planes_of(buf, n, width);
  float* meta = new_meta(st, buf);
  NF_REQUIRE(meta, "tcgen05 path: out of tensor metadata slots");
  *out = TcPlanesOut{P.p, P.plane_elems(), P.ld, meta};
  return NF_OK;
}

float* tc_alloc_meta(Flow& f) {
  TcState* st = get_state(f);
  if (!st || st->next_slot >= kMetaSlots) return nullptr;
  return st->meta_pool + 2 * (st->next_slot++);
}

int tc_absmax(Flow& f, const float* X, int64_t count, float* meta) {
  absmax_kernel<<<(unsigned)std::min<int64_t>(ceil_div(count, 256 * 8), 4 * kNumSMs), 256, 0, f.stream>>>(X, 1, nullptr, 1, count, meta);
  NF_LAUNCH_CHECK();
  return NF_OK;
}

int tc_mlp_forward(Flow& f, const LayerDesc& Ld, int m, int64_t n, void* act0, std::vector<void*>& acts) {
  TcState* st = get_state(f);
  const int li = (int)(&Ld - f.layers.data());
  const MLPDesc& md = Ld.mlps[m];
  const int terms = f.mma_mode == NF_MMA_F16X1 ? 1 : 3;
  const int nd = md.n_dense();
  for (int i = 0; i < nd; ++i) {
    const int pi = st->index[li][m][i];
    const DensePrep& dp = st->preps[pi];
    const bool last = (i + 1 == nd);
    void* abuf = i == 0 ? act0 : acts[i - 1];
    Planes A = planes_of(abuf, n, dp.kin);
    // a one-chunk contraction (K <= 64) is all epilogue: 64-column tiles run two CTAs per SM whose phases interleave
    static const int fwd_bn_k64 = getenv("NFCUDA_FWD_BN_K64") ? atoi(getenv("NFCUDA_FWD_BN_K64")) : 256;
    int bn = std::min(dp.nf_rows, 256);
    if (dp.kin_p == 64 && dp.nf_rows % fwd_bn_k64 == 0 && fwd_bn_k64 < bn) bn = fwd_bn_k64;
    const int n_tiles_n = dp.nf_rows / bn;
    CUtensorMap ma, mb;
    NF_TRY(make_map_kmajor(st, A.p, n, A.ld, A.plane_elems(), 128, &ma));
    NF_TRY(make_map_kmajor(st, st->pool + dp.wf_off, dp.nf_rows, dp.kin_p, (int64_t)dp.nf_rows * dp.kin_p, bn, &mb));
    GemmParams p{};
    p.M = n; p.num_k_chunks = dp.kin_p / 64; p.terms = terms;
    p.a_meta = meta_of(st, abuf);
    NF_REQUIRE(p.a_meta, "tcgen05 path: missing tensor metadata (forward input)");
    p.w_sc = st->d_scalars + 4 * pi;
    p.bound_dgrad = 0;
    p.slab_stages = slab_stages_for(false, p.num_k_chunks);
    p.rz_comp = rz_compensation(dp.kin, p.num_k_chunks, p.slab_stages);
    p.bias = st->bias_pool + dp.bias_off;
    CUtensorMap mo = ma;
    if (!last) {
      Planes O = planes_of(acts[i], n, dp.nout);
      NF_TRY(make_map_store(st, O.p, n, O.ld, O.plane_elems(), &mo));
      p.out_meta = new_meta(st, acts[i]);
      NF_REQUIRE(p.out_meta, "tcgen05 path: out of tensor metadata slots");
      p.epi = EPI_PLANES_ACT; p.act = ACT_LRELU; p.n_store = O.ld;
      p.out_hi = O.p; p.out_lo = O.p + O.plane_elems(); p.out_ld = O.ld;
      p.out_bits = O.bits(); p.out_bits_ld = O.bits_ld();
    } else {
      p.epi = EPI_F32_ACT; p.act = md.out_act ? ACT_TANH : ACT_NONE; p.n_store = dp.nout;
      p.out_f32 = (float*)acts[i]; p.out_f32_ld = dp.nout;
      static const bool f32_tma_ok = !(getenv("NFCUDA_F32_TMA_STORE") && atoi(getenv("NFCUDA_F32_TMA_STORE")) == 0);
      if (f32_tma_ok && (dp.nout & 3) == 0 && (reinterpret_cast<uintptr_t>(acts[i]) & 15) == 0) {
        NF_TRY(make_map_store_f32(st, acts[i], n, dp.nout, dp.nout, &mo));
        p.f32_tma = 1;
      }
    }
    NF_TRY(launch_gemm(f, bn, ma, mb, mo, p, n_tiles_n));
  }
  return NF_OK;
}

int tc_mlp_backward(Flow& f, const LayerDesc& Ld, int m, int64_t n, void* act0, std::vector<void*>& acts, float* g_last,
                    const float* g_last_amax, void* scratch0, void* scratch1, float* G, double* gsum, bool last_bias_done) {
  TcState* st = get_state(f);
  const int li = (int)(&Ld - f.layers.data());
  const MLPDesc& md = Ld.mlps[m];
  const int terms = f.mma_mode == NF_MMA_F16X1 ? 1 : 3;
  const int nd = md.n_dense();
  void* gbuf = scratch0;
  void* gnext = scratch1;
  // Bias gradients as column sums of G inside the weight-gradient kernel (its epilogue warps idle during the K loop) instead of
  // in the dgrad epilogue: OFF by default.  Measured (C3, 2^20): the dgrad GEMMs gain 1.4 ms per step (K=256: 0.446 -> 0.406 ms,
  // K=64: 0.320 -> 0.271 ms) but the weight-gradient kernels lose 3-4 ms (256x256: 0.393 -> 0.490 ms) -- their MMAs read both
  // MN-major operands from shared memory at ~75 % of its bandwidth, and the extra 32 KB of reads per stage do not fit.
  static const bool wgrad_colsum = getenv("NFCUDA_WGRAD_COLSUM") && atoi(getenv("NFCUDA_WGRAD_COLSUM")) != 0;
  // gradient w.r.t. the last pre-activation -> split planes
  // ... and, in the same pass, the bias gradient of the last Dense (column sums of the fp32 gradient)
  if (g_last)
    NF_TRY(split_into_planes(f, st, g_last, md.dims[nd], nullptr, md.dims[nd], n, gbuf, g_last_amax,
                             last_bias_done ? nullptr : gsum + md.b_off[nd - 1]));
  else NF_REQUIRE(last_bias_done && meta_of(st, gbuf), "tc_mlp_backward: planes input without metadata / bias gradient");
  for (int i = nd - 1; i >= 0; --i) {
    const int pi = st->index[li][m][i];
    const DensePrep& dp = st->preps[pi];
    Planes Gp = planes_of(gbuf, n, dp.nout);
    void* xbuf = i == 0 ? act0 : acts[i - 1];
    Planes X = planes_of(xbuf, n, dp.kin);
    const float* g_meta = meta_of(st, gbuf);
    const float* x_meta = meta_of(st, xbuf);
    NF_REQUIRE(g_meta && x_meta, "tcgen05 path: missing tensor metadata (backward)");
    {  // weight gradient
      // the kernel holds a [<=256 in] x [<=256 out] block of dW^T in TMEM; wider layers are covered block by block
      for (int i0 = 0; i0 < X.ld; i0 += 256) {
        for (int j0 = 0; j0 < Gp.ld; j0 += 256) {
          const int xl = std::min(256, X.ld - i0), gl = std::min(256, Gp.ld - j0);
          if (i0 >= dp.kin || j0 >= dp.nout) continue;
          CUtensorMap mx, mg;
          const int mt = (int)ceil_div(xl, 128);
          NF_TRY(make_map_mnmajor(st, X.p + i0, n, X.ld, X.plane_elems(), mt * 2, &mx, xl));
          NF_TRY(make_map_mnmajor(st, Gp.p + j0, n, Gp.ld, Gp.plane_elems(), std::max(1, gl / 64), &mg, gl));
          WgradParams wp{};
          wp.n = n; wp.kin = std::min(256, dp.kin - i0); wp.nout = std::min(256, dp.nout - j0); wp.mt = mt; wp.terms = terms;
          wp.gW = gsum + dp.w_off + (int64_t)i0 * dp.nout + j0; wp.ldw = dp.nout;
          wp.x_meta = x_meta; wp.g_meta = g_meta;
          // bias gradient of this Dense = column sums of its pre-activation gradient G: summed here from the G stages (idle
          // epilogue warps) instead of in the epilogue of the dgrad GEMM that produced G; the last Dense's comes from
          // split_into_planes above
          wp.colsum = (wgrad_colsum && i0 == 0 && i < nd - 1) ? gsum + dp.b_off + j0 : nullptr;
          switch (gl) {
            case 32: case 64: NF_TRY(launch_wgrad_bn<64>(f, mx, mg, wp)); break;
            case 128: NF_TRY(launch_wgrad_bn<128>(f, mx, mg, wp)); break;
            case 192: NF_TRY(launch_wgrad_bn<192>(f, mx, mg, wp)); break;
            default: NF_TRY(launch_wgrad_bn<256>(f, mx, mg, wp)); break;
          }
        }
      }
    }
    if (i > 0 || G) {  // data gradient (skipped for the first Dense when nobody needs d/d(conditioner input))
      // a one-chunk contraction (K <= 64) is all epilogue: narrower tiles run two CTAs per SM whose phases interleave
      static const int dgrad_bn_k64 = getenv("NFCUDA_DGRAD_BN_K64") ? atoi(getenv("NFCUDA_DGRAD_BN_K64")) : 64;   // measured (C3, 2^20): 256: 0.388 ms, 128: 0.350, 64: 0.324, 32: 0.487
      int bn = std::min(dp.nd_rows, 256);
      if (dp.nout_p == 64 && dp.nd_rows % dgrad_bn_k64 == 0 && dgrad_bn_k64 < bn) bn = dgrad_bn_k64;
      static const int dgrad_bn_big = getenv("NFCUDA_DGRAD_BN_BIG") ? atoi(getenv("NFCUDA_DGRAD_BN_BIG")) : 256;
      if (dp.nout_p > 64 && dp.nd_rows % dgrad_bn_big == 0 && dgrad_bn_big < bn) bn = dgrad_bn_big;
      const int n_tiles_n = dp.nd_rows / bn;
      CUtensorMap ma, mb;
      NF_TRY(make_map_kmajor(st, Gp.p, n, Gp.ld, Gp.plane_elems(), 128, &ma));
      NF_TRY(make_map_kmajor(st, st->pool + dp.wd_off, dp.nd_rows, dp.nout_p, (int64_t)dp.nd_rows * dp.nout_p, bn, &mb));
      GemmParams p{};
      p.M = n; p.num_k_chunks = dp.nout_p / 64; p.terms = terms;
      p.a_meta = g_meta; p.w_sc = st->d_scalars + 4 * pi; p.bound_dgrad = 1;
      p.slab_stages = slab_stages_for(true, p.num_k_chunks);
      p.rz_comp = rz_compensation(dp.nout, p.num_k_chunks, p.slab_stages);
      CUtensorMap mo = ma;
      if (i > 0) {
        Planes O = planes_of(gnext, n, dp.kin);
        NF_TRY(make_map_store(st, O.p, n, O.ld, O.plane_elems(), &mo));
        p.out_meta = new_meta(st, gnext);
        NF_REQUIRE(p.out_meta, "tcgen05 path: out of tensor metadata slots");
        p.epi = EPI_PLANES_MASK; p.n_store = O.ld;
        p.out_hi = O.p; p.out_lo = O.p + O.plane_elems(); p.out_ld = O.ld;
        p.mask_bits = X.bits(); p.mask_ld = X.bits_ld();
        const DensePrep& below = st->preps[st->index[li][m][i - 1]];
        if (!wgrad_colsum) { p.colsum_out = gsum + below.b_off; p.colsum_n = below.nout; }   // bias gradient of Dense i-1, fused (A/B path)
      } else {
        p.epi = EPI_SCATTER_ADD; p.n_store = dp.kin; p.G = G; p.ldg = f.dim; p.idx = Ld.d_idx2;
      }
      NF_TRY(launch_gemm(f, bn, ma, mb, mo, p, n_tiles_n));
      std::swap(gbuf, gnext);
    }
  }
  return NF_OK;
}


// ---------------------------------------------------------------------------------------------
// fused AffineCoupling forward (fused_coupling.cuh)
// ---------------------------------------------------------------------------------------------
bool tc_fused_affine_ok(const Flow& f, const LayerDesc& Ld) {
  if (!g_opt_fused_coupling || f.dtype != NF_F32 || f.mma_mode == NF_MMA_SIMT) return false;
  if (Ld.kind != NF_AFFINE_COUPLING || Ld.mlps.size() != 2) return false;
  const int d = f.dim, c = (int)Ld.idx1.size(), cbar = (int)Ld.idx2.size();
  if ((d & 3) || d > 64 || c < 1 || c > 32 || cbar < 1 || cbar > 64) return false;
  for (const MLPDesc& md : Ld.mlps) {
    if (md.n_dense() != 3 || md.dims[0] != cbar || md.dims[3] != c) return false;
    // hidden layers of at most 32 columns keep narrow planes (plane_ld), which the fused kernel's stash stores do not write
    if (md.dims[1] != md.dims[2] || pad64(md.dims[1]) > 256 || plane_ld(md.dims[1]) != pad64(md.dims[1])) return false;
  }
  return Ld.mlps[0].dims[1] == Ld.mlps[1].dims[1];
}

int tc_affine_forward_fused(Flow& f, const LayerDesc& Ld, int64_t n, const float* Xin, float* Xout, float* ld, void* act0,
                            std::vector<std::vector<void*>>& acts, const float* x_meta, float* y_meta, bool inv, bool stash) {
  TcState* st = get_state(f);
  NF_REQUIRE(st && x_meta, "fused coupling: missing state / input bound");
  const int li = (int)(&Ld - f.layers.data());
  const int d = f.dim, c = (int)Ld.idx1.size(), cbar = (int)Ld.idx2.size();
  const int H = Ld.mlps[0].dims[1], h_ld = pad64(H);
  int*& d_pos2 = st->pos2[li];
  if (!d_pos2) {
    std::vector<int> pos2(d, -1);
    for (int k = 0; k < cbar; ++k) pos2[Ld.idx2[k]] = k;
    NF_CUDA(cudaMalloc((void**)&d_pos2, d * sizeof(int)));
    NF_CUDA(cudaMemcpy(d_pos2, pos2.data(), d * sizeof(int), cudaMemcpyHostToDevice));
  }
  FusedFwdMaps maps;
  FusedFwdParams p{};
  p.n = n; p.d = d; p.c = c; p.cbar = cbar; p.nch = h_ld / 64; p.kk1 = (cbar + 15) / 16; p.inv = inv ? 1 : 0;
  p.terms = f.mma_mode == NF_MMA_F16X1 ? 1 : 3;
  p.Xin = Xin; p.Xout = Xout; p.ld = ld; p.pos = Ld.d_pos; p.pos2 = d_pos2; p.x_meta = x_meta; p.y_meta = y_meta;
  p.h_ld = h_ld; p.h_plane_elems = round_up(n, 128) * h_ld;
  Planes A0 = planes_of(act0, n, cbar);
  NF_REQUIRE(A0.ld == 64 || A0.ld == 32, "fused coupling: conditioner input wider than 64");
  NF_TRY(make_map_kmajor(st, A0.p, n, A0.ld, A0.plane_elems(), 128, &maps.x2));
  p.x2_meta = new_meta(st, act0);
  NF_REQUIRE(p.x2_meta, "tcgen05 path: out of tensor metadata slots");
  for (int m = 0; m < 2; ++m) {
    FusedNet& N = p.net[m];
    for (int i = 0; i < 3; ++i) {
      const int pi = st->index[li][m][i];
      const DensePrep& dp = st->preps[pi];
      N.w_sc[i] = st->d_scalars + 4 * pi;
      N.bias[i] = st->bias_pool + dp.bias_off;
      NF_REQUIRE(dp.kin_p == (i == 0 ? 64 : h_ld) && dp.nf_rows >= (i == 2 ? 32 : h_ld), "fused coupling: unexpected weight plane shape");
      NF_TRY(make_map_kmajor(st, st->pool + dp.wf_off, dp.nf_rows, dp.kin_p, (int64_t)dp.nf_rows * dp.kin_p, i == 2 ? 32 : 64, &maps.w[m][i]));
      if (i < 2) NF_TRY(make_map_kmajor(st, st->pool + dp.wf_off, dp.nf_rows, dp.kin_p, (int64_t)dp.nf_rows * dp.kin_p, 128, &maps.w128[m][i]));
      if (i < 2) {
        Planes O = planes_of(acts[m][i], n, H);
        N.h_planes[i] = O.p;
        NF_TRY(make_map_store32(st, O.p, n, O.ld, O.plane_elems(), &maps.h[m][i]));
        NF_TRY(make_map_store64(st, O.p, n, O.ld, O.plane_elems(), &maps.h64[m][i]));
        N.h_bits[i] = reinterpret_cast<uint16_t*>(O.bits());
        N.h_meta[i] = new_meta(st, acts[m][i]);
        NF_REQUIRE(N.h_meta[i], "tcgen05 path: out of tensor metadata slots");
      }
    }
    N.out = (m == 0 && stash) ? (float*)acts[m][2] : nullptr;      // the backward pass needs s (after tanh) only
  }
  p.dbg_flags = getenv("NFCUDA_DBG_FLAGS") ? atoi(getenv("NFCUDA_DBG_FLAGS")) : 0;
  p.no_stash = stash ? 0 : 1;      // sampling / logpdf / plain transforms: nothing will read the stash
  if (!stash) p.dbg_flags |= 1;
  {
    const int dm = (p.dbg_flags >> 8) & 15, de = (p.dbg_flags >> 12) & 15;
    static const int slab_env = getenv("NFCUDA_FUSED_SLAB") ? atoi(getenv("NFCUDA_FUSED_SLAB")) : 2;
    p.slab = slab_env >= 2 ? 2 : 1;
    static const int hoist_env = getenv("NFCUDA_FUSED_HOIST") ? atoi(getenv("NFCUDA_FUSED_HOIST")) : 4;   // measured: all chunks hoisted 0.271 ms, two 0.277 ms (C3, 2^17)
    p.n_hoist = std::min(h_ld / 64, std::max(hoist_env, p.slab));     // the first chains read chunks 0 .. slab - 1
    // Third-Dense placement: `delay` items after the chunk's last K chunk in the issuer's copy (measured, C3 2^17: 2: 0.259 ms,
    // 3: 0.263, 4: 0.266), later in the epilogue's copy (4 / 6: same, 9: +3 %).  The issuer's delay must not exceed one turn of
    // each team (2 x slab items): beyond that the slab sits behind a chain of the OTHER team, whose accumulator that team only
    // drains after its own activation pass -- which waits for this very slab (h2_free): a cycle (delay 6 hangs).
    const int dm_eff = std::min(dm ? dm : 2, 2 * p.slab), de_eff = std::min(de ? de : 6, 9);
    const int lead = (p.dbg_flags >> 20) & 1 ? 2 : 1;      // experiment: team 0 two chains ahead instead of one
    p.n_seq_m = fused_build_schedule(p.seq_m, (int)sizeof(p.seq_m), h_ld / 64, dm_eff, p.slab, p.n_hoist, lead);
    p.n_seq_e = fused_build_schedule(p.seq_e, (int)sizeof(p.seq_e), h_ld / 64, de_eff, p.slab, p.n_hoist, lead);
    NF_REQUIRE(p.n_seq_m <= (int)sizeof(p.seq_m) && p.n_seq_e == p.n_seq_m, "fused coupling: schedule does not fit");
  }
  p.rz[0] = rz_compensation(cbar, 1, 1);
  // two-K-chunk accumulation chains of the two-team kernel: calibrated on C3-shaped conditioners (tests/tools/fused_check.py;
  // 0.8e-8 .. 3.2e-8 keep the ELBO within 4e-6 of the Float64 oracle at four couplings, 2.4e-8 gives 1.2e-6)
  static const float c2_fused = getenv("NFCUDA_RZ_C2F") ? (float)atof(getenv("NFCUDA_RZ_C2F")) : 2.4e-8f;
  p.rz[1] = rz_compensation(H, h_ld / 64, g_opt_fused_variant == 0 ? p.slab : 1, c2_fused);
  p.rz[2] = rz_compensation(H, h_ld / 64, 1);
  const bool wide = g_opt_fused_variant != 0;
  static bool attr_set[64] = {};
  if (!attr_set[f.device & 63]) {
    NF_CUDA(cudaFuncSetAttribute(fused_affine_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedCfg::SMEM));
    NF_CUDA(cudaFuncSetAttribute(fused_affine_fwd_w128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F2Cfg::SMEM));
    attr_set[f.device & 63] = true;
  }
  const int64_t tiles = ceil_div(n, 128);
  const unsigned grid = (unsigned)std::min<int64_t>(tiles, kNumSMs);
  long long* d_dbg = nullptr;
  static int dbg_seen = 0;
  if (getenv("NFCUDA_DBG") && !strcmp(getenv("NFCUDA_DBG"), "fused_fwd") && dbg_seen++ == (getenv("NFCUDA_DBG_SKIP") ? atoi(getenv("NFCUDA_DBG_SKIP")) : 0)) {
    NF_CUDA(cudaMalloc((void**)&d_dbg, 5 * 512 * sizeof(long long)));
    NF_CUDA(cudaMemset(d_dbg, 0, 5 * 512 * sizeof(long long)));
    p.dbg = d_dbg;
  }
  f.prof.begin("fused_affine_fwd", f.stream);
  if (wide) fused_affine_fwd_w128_kernel<<<grid, F2Cfg::THREADS, F2Cfg::SMEM, f.stream>>>(maps, p);
  else fused_affine_fwd_kernel<<<grid, FusedCfg::THREADS, FusedCfg::SMEM, f.stream>>>(maps, p);
  f.prof.end(f.stream);
  NF_LAUNCH_CHECK();
  if (d_dbg) {
    std::vector<long long> h(5 * 512);
    NF_CUDA(cudaStreamSynchronize(f.stream));
    NF_CUDA(cudaMemcpy(h.data(), d_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(d_dbg);
    long long t0 = 0;
    for (auto v : h) if (v && (!t0 || v < t0)) t0 = v;
    fprintf(stderr, "[nfcuda dbg] fused_fwd grid %u tiles %lld: clock64 relative to the first event (CTA 0)\n", grid, (long long)tiles);
    fprintf(stderr, "  schedule (MMA):");
    for (int i = 0; i < p.n_seq_m; ++i) fprintf(stderr, " %02x", p.seq_m[i]);
    fprintf(stderr, "\n  schedule (epilogue):");
    for (int i = 0; i < p.n_seq_e; ++i) fprintf(stderr, " %02x", p.seq_e[i]);
    fprintf(stderr, "\n");
    if (wide) {     // tagged event list of epilogue thread 0: (tag, cycles since the previous event)
      long long prev = h[512];
      fprintf(stderr, "  events:");
      for (int i = 0; i < 511 && h[512 + i]; ++i) { fprintf(stderr, " %lld:+%lld", h[1024 + i], h[512 + i] - prev); prev = h[512 + i]; }
      fprintf(stderr, "\n");
    } else
    for (int r = 0; r < 5; ++r) {
      fprintf(stderr, "  role %d:", r);
      for (int i = 0; i < 512; ++i) { if (i % 4 == 0) fprintf(stderr, " |"); fprintf(stderr, " %lld", h[r * 512 + i] ? h[r * 512 + i] - t0 : -1); }
      fprintf(stderr, "\n");
    }
  }
  return NF_OK;
}

int tc_fused_schedule(int nch, int slab, int n_hoist, int delay, unsigned char* out, int cap) {
  return fused_build_schedule(out, cap, nch, delay, slab, n_hoist);
}

// Test hook: Y[n, N] = X[n, K] * Wt[K, N] + b through the tcgen05 forward GEMM (one Dense, no activation).
int tc_gemm_selftest(int64_t n, int K, int N, const float* X_host, const float* Wt_host, const float* b_host, int terms,
                     float* Y_host) {
  Flow f;
  f.dim = K + N; f.dtype = NF_F32;
  f.mma_mode = terms == 1 ? NF_MMA_F16X1 : NF_MMA_F16X3;
  NF_CUDA(cudaGetDevice(&f.device));
  NF_CUDA(cudaStreamCreateWithFlags(&f.stream, cudaStreamNonBlocking));
  LayerDesc L;
  L.kind = NF_AFFINE_COUPLING;
  MLPDesc md;
  md.dims = {K, N};
  md.w_off = {0}; md.b_off = {(int64_t)K * N};
  md.out_act = 0;
  L.mlps.push_back(md);
  f.layers.push_back(L);
  f.P = (int64_t)K * N + N;
  float *theta = nullptr, *X = nullptr, *Y = nullptr;
  void* act0 = nullptr;
  int status = NF_OK;
  auto body = [&]() -> int {
    NF_CUDA(cudaMalloc((void**)&theta, f.P * sizeof(float)));
    NF_CUDA(cudaMalloc((void**)&X, (size_t)n * K * sizeof(float)));
    NF_CUDA(cudaMalloc((void**)&Y, (size_t)n * N * sizeof(float)));
    NF_CUDA(cudaMalloc(&act0, tc_act_bytes(n, K)));
    NF_CUDA(cudaMemcpy(theta, Wt_host, (size_t)K * N * sizeof(float), cudaMemcpyHostToDevice));
    NF_CUDA(cudaMemcpy(theta + (size_t)K * N, b_host, (size_t)N * sizeof(float), cudaMemcpyHostToDevice));
    NF_CUDA(cudaMemcpy(X, X_host, (size_t)n * K * sizeof(float), cudaMemcpyHostToDevice));
    NF_TRY(tc_prepare_weights(f, theta));
    NF_TRY(tc_gather_split(f, X, K, nullptr, K, n, act0, nullptr));
    std::vector<void*> acts{(void*)Y};
    NF_TRY(tc_mlp_forward(f, f.layers[0], 0, n, act0, acts));
    NF_CUDA(cudaStreamSynchronize(f.stream));
    NF_CUDA(cudaMemcpy(Y_host, Y, (size_t)n * N * sizeof(float), cudaMemcpyDeviceToHost));
    return NF_OK;
  };
  status = body();
  cudaFree(theta); cudaFree(X); cudaFree(Y); cudaFree(act0);
  tc_release(f);
  cudaStreamDestroy(f.stream);
  return status;
}

}  // namespace nf
