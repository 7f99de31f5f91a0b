// K3f, wide variant -- the fused AffineCoupling forward of fused_coupling.cuh with 128-column MMAs.
//
// Measured on B200 (profiles/r2_fused_fwd_notes.md): a tcgen05.mma whose A operand lives in tensor memory re-reads the
// 128 x 16 A tile (4 KB) for every instruction at ~64 B/clk, so a 128 x N x 16 MMA costs max(N/2, 64) cycles -- the
// 64-column MMAs of the first version run at half rate.  Here every hidden-layer MMA is 128 x 128 x 16:
//
//   L1: acc[128 x 128] = x2 W1[chunk]^T       (A from shared memory)      2 chunks per network
//   L2: acc[128 x 128] = h1 W2[chunk, slab]^T  (A from tensor memory)      2 chunks x 4 K slabs, slabs summed in registers
//   L3: acc3[128 x 32] += h2[sub-chunk of 64] W3[:, sub-chunk]^T           4 sub-chunks, through a 64-column operand buffer
//
// TMEM (512 columns): [0,128) h1 hi | [128,256) h1 lo | [256,384) ONE 128-column accumulator |
//                     [384,416) h2 sub-chunk hi | [416,448) h2 sub-chunk lo | [448,512) two 32-column L3 accumulators
// One accumulator only (the budget is 512 columns): the drain of slab i and the issue of slab i+1 serialise, so the schedule
// interleaves the L3 MMAs of the previous chunk with the first slabs of the next one, and each epilogue thread owns 16 columns
// of EACH 64-column half of a chunk -- the whole warp set finishes half a first (L3 can start) while it still works on half b.
// Weight slabs are 32 KB (128 rows x 64 K, hi + lo) in a 3-stage ring; the X tile has one buffer (the coupling arithmetic
// re-reads its row from global memory), so the next tile's X arrives while this tile computes.
//
// Same parameters, stash formats and results as fused_affine_fwd_kernel.  Included by tc_gemm.cu after fused_coupling.cuh.

struct F2Cfg {
  static constexpr int STAGES = 3;
  static constexpr int STAGE = 32768;            // weight slab: 128 rows x 128 B hi plane, then the lo plane (+16384)
  static constexpr int XS = 32768;
  static constexpr int X2_PLANE = 16384;
  static constexpr int ST_LD = 33;
  static constexpr int OFF_W = 0;
  static constexpr int OFF_X = OFF_W + STAGES * STAGE;
  static constexpr int OFF_X2 = OFF_X + XS;
  static constexpr int OFF_S = OFF_X2 + 2 * X2_PLANE;
  static constexpr int OFF_STG = OFF_S + 128 * ST_LD * 4;         // 16 warps x 2 KB stash staging; the t tile aliases it
  static constexpr int OFF_T = OFF_STG;
  static constexpr int OFF_BITS = OFF_STG + 16 * 2048;            // [2 layers][128 rows][h_ld / 16] uint16
  static constexpr int OFF_BIAS = OFF_BITS + 2 * 4096;
  static constexpr int OFF_LD = OFF_BIAS + 2 * 544 * 4;
  static constexpr int OFF_BAR = OFF_LD + 512;
  static constexpr int N_BARS = 2 * STAGES + 2 /*x*/ + 2 /*x2*/ + 2 /*acc*/ + 4 /*acc3*/ + 4 /*h1 ready per K slab*/ + 2 /*h2*/;
  static constexpr int OFF_SEQ = OFF_BAR + 8 * N_BARS + 16;
  static constexpr int OFF_POS = OFF_SEQ + 64;                    // pos[64], pos2[64]: the masks, read per element by the scatter / coupling passes
  static constexpr int SMEM = OFF_POS + 512;
  static constexpr int EPI0 = 128;
  static constexpr int THREADS = EPI0 + 512;
  static constexpr int TM_H1HI = 0, TM_H1LO = 128, TM_ACC = 256, TM_H2HI = 384, TM_H2LO = 416, TM_ACC3 = 448;
};
static_assert(F2Cfg::SMEM <= 232448, "fused coupling (wide): shared memory budget");

// schedule after the first Dense: L2 slab (j, k) = (j << 2) | k ; L3 sub-chunk q = 0x80 | q.  The L3 items of chunk j - 1
// follow the FIRST slab of chunk j (their operand is ready by then and they fill the drain gap of the single accumulator).
__device__ __forceinline__ int f2_build_seq(uint8_t* seq, int nck, int nks) {
  int n = 0;
  for (int j = 0; j < nck; ++j)
    for (int k = 0; k < nks; ++k) {
      seq[n++] = (uint8_t)((j << 2) | k);
      if (j > 0 && k == 0)
        for (int q = 2 * (j - 1); q < 2 * j && q < nks; ++q) seq[n++] = (uint8_t)(0x80 | q);
    }
  for (int q = 2 * (nck - 1); q < nks; ++q) seq[n++] = (uint8_t)(0x80 | q);
  return n;
}

// bias + leakyrelu + sign bits + hi/lo split of 16 scaled pre-activations
__device__ __forceinline__ uint32_t f2_act_split(const float (&acc)[16], float ds, const float* __restrict__ bias, uint32_t (&hi)[8],
                                                 uint32_t (&lo)[8]) {
  uint32_t bits = 0;
#pragma unroll
  for (int q = 0; q < 16; q += 2) {
    float a = fmaf(acc[q], ds, bias[q]);
    float b = fmaf(acc[q + 1], ds, bias[q + 1]);
    bits |= (a > 0.f ? 1u : 0u) << q;
    bits |= (b > 0.f ? 1u : 0u) << (q + 1);
    a = fmaxf(a, 0.01f * a); b = fmaxf(b, 0.01f * b);
    split_pair(a, b, hi[q >> 1], lo[q >> 1]);
  }
  return bits;
}

__global__ void __launch_bounds__(F2Cfg::THREADS, 1)
fused_affine_fwd_w128_kernel(const __grid_constant__ FusedFwdMaps maps, const FusedFwdParams p) {
  using C = F2Cfg;
  constexpr int S = C::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  const uint32_t bars = base + C::OFF_BAR;
  auto w_full = [&](int s) { return bars + 8u * s; };
  auto w_empty = [&](int s) { return bars + 8u * (S + s); };
  const uint32_t x_full = bars + 8u * (2 * S), x_empty = bars + 8u * (2 * S + 1);
  const uint32_t x2_ready = bars + 8u * (2 * S + 2), x2_free = bars + 8u * (2 * S + 3);
  const uint32_t tfull = bars + 8u * (2 * S + 4), tempty = bars + 8u * (2 * S + 5);
  auto tfull3 = [&](int a) { return bars + 8u * (2 * S + 6 + a); };
  auto tempty3 = [&](int a) { return bars + 8u * (2 * S + 8 + a); };
  auto h1_ready = [&](int k) { return bars + 8u * (2 * S + 10 + k); };
  const uint32_t h2_ready = bars + 8u * (2 * S + 14), h2_free = bars + 8u * (2 * S + 15);
  const uint32_t tmem_slot = bars + 8u * C::N_BARS;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + C::OFF_BAR + 8 * C::N_BARS);
  uint8_t* seq = smem_raw + C::OFF_SEQ;
  float* s_bias = reinterpret_cast<float*>(smem_raw + C::OFF_BIAS);
  float* s_ld = reinterpret_cast<float*>(smem_raw + C::OFF_LD);
  float* s_S = reinterpret_cast<float*>(smem_raw + C::OFF_S);
  float* s_T = reinterpret_cast<float*>(smem_raw + C::OFF_T);
  int* s_pos = reinterpret_cast<int*>(smem_raw + C::OFF_POS);
  int* s_pos2 = s_pos + 64;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int64_t num_tiles = (p.n + 127) / 128;
  const int nks = p.nch;                       // 64-wide K slabs of the hidden width (1..4)
  const int nck = (nks + 1) >> 1;              // 128-wide output chunks
  const int n_seq = nck * nks + nks;

  const float amax_x = __uint_as_float(reinterpret_cast<const unsigned int*>(p.x_meta)[1]);
  const float s_x2 = pow2_scale(amax_x);

  if (warp == 0 && lane == 0) {
    for (int nt = 0; nt < 2; ++nt) { tma_prefetch_desc(&maps.w128[nt][0]); tma_prefetch_desc(&maps.w128[nt][1]); tma_prefetch_desc(&maps.w[nt][2]); }
    tma_prefetch_desc(&maps.x2);
    for (int s = 0; s < S; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    mbar_init(x_full, 1); mbar_init(x_empty, 16);
    mbar_init(x2_ready, 16); mbar_init(x2_free, 1);
    mbar_init(tfull, 1); mbar_init(tempty, 16);
    for (int a = 0; a < 2; ++a) { mbar_init(tfull3(a), 1); mbar_init(tempty3(a), 16); }
    for (int k = 0; k < 4; ++k) mbar_init(h1_ready(k), 16);
    mbar_init(h2_ready, 16); mbar_init(h2_free, 1);
    fence_barrier_init();
    f2_build_seq(seq, nck, nks);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (threadIdx.x >= C::EPI0) {
    const int t = threadIdx.x - C::EPI0;
    for (int i = t; i < 2 * C::X2_PLANE / 16; i += 512) reinterpret_cast<uint4*>(smem_raw + C::OFF_X2)[i] = make_uint4(0, 0, 0, 0);
    if (t < 128) s_ld[t] = 0.f;
    if (t < p.d) { s_pos[t] = p.pos[t]; s_pos2[t] = p.pos2[t]; }
    for (int i = t; i < 2 * 544; i += 512) {
      const int nt = i / 544, o = i % 544;
      const float bound1 = amax_x * p.net[nt].w_sc[0][1] + p.net[nt].w_sc[0][3];
      const float bound2 = bound1 * p.net[nt].w_sc[1][1] + p.net[nt].w_sc[1][3];
      float v;
      if (o < 256) v = o < 64 * nks ? p.net[nt].bias[0][o] * pow2_scale(bound1 * 1.001f) : 0.f;
      else if (o < 512) v = (o - 256) < 64 * nks ? p.net[nt].bias[1][o - 256] * pow2_scale(bound2 * 1.001f) : 0.f;
      else v = p.net[nt].bias[2][o - 512];
      s_bias[i] = v;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // =========================== producer ===========================
    uint32_t it = 0, xt = 0;
    auto load_x = [&](int64_t tile) {
      mbar_wait(x_empty, (xt & 1) ^ 1);
      const int64_t r0 = tile * 128;
      const int rows = (int)((p.n - r0) < 128 ? (p.n - r0) : 128);
      const uint32_t bytes = (uint32_t)rows * (uint32_t)p.d * 4u;
      if (elect_one_sync()) {
        mbar_expect_tx(x_full, bytes);
        bulk_load_1d(base + C::OFF_X, p.Xin + r0 * p.d, bytes, x_full);
      }
      __syncwarp();
      ++xt;
    };
    auto load_w = [&](const CUtensorMap* map, int c0, int c1, uint32_t bytes_per_plane) {
      const int s = it % S;
      mbar_wait(w_empty(s), ((it / S) & 1) ^ 1);
      const uint32_t st = base + C::OFF_W + s * C::STAGE;
      if (elect_one_sync()) {
        mbar_expect_tx(w_full(s), (p.terms > 1 ? 2u : 1u) * bytes_per_plane);
        tma_load_3d(st, map, w_full(s), c0, c1, 0);
        if (p.terms > 1) tma_load_3d(st + 16384, map, w_full(s), c0, c1, 1);
      }
      __syncwarp();
      ++it;
    };
    if ((int64_t)blockIdx.x < num_tiles) load_x(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int nt = 0; nt < 2; ++nt) {
        for (int j = 0; j < nck; ++j) load_w(&maps.w128[nt][0], 0, j * 128, 16384);
        // the X buffer is released as soon as this tile's conditioner input has been built: the next tile's rows land early
        if (nt == 0 && tile + gridDim.x < num_tiles) load_x(tile + gridDim.x);
        for (int i = 0; i < n_seq; ++i) {
          const int e = seq[i];
          if (e & 0x80) load_w(&maps.w[nt][2], (e & 3) * 64, 0, 4096);
          else load_w(&maps.w128[nt][1], (e & 3) * 64, ((e >> 2) & 3) * 128, 16384);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc128 = make_idesc(128, 128, 0, 0);
    constexpr uint32_t idesc32 = make_idesc(128, 32, 0, 0);
    uint32_t it = 0, sl = 0, sl3 = 0, tcount = 0, ncount = 0, cc = 0;
    const bool t3 = p.terms > 1;
    const uint32_t x2a = base + C::OFF_X2;
    const int kk1 = p.kk1;
    const uint32_t d_acc = tmem_base + C::TM_ACC;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
      mbar_wait(x2_ready, tcount & 1);
      tc_fence_after();
      for (int nt = 0; nt < 2; ++nt, ++ncount) {
        // ---- first Dense: A = x2 planes in shared memory ----
        for (int j = 0; j < nck; ++j, ++it, ++sl) {
          const int s = it % S;
          mbar_wait3(tempty, (sl & 1) ^ 1, w_full(s), (it / S) & 1, 0, 0);
          tc_fence_after();
          const uint32_t st = base + C::OFF_W + s * C::STAGE;
          const uint64_t a_hi = make_smem_desc(x2a, 16, 1024), a_lo = make_smem_desc(x2a + C::X2_PLANE, 16, 1024);
          const uint64_t b_hi = make_smem_desc(st, 16, 1024), b_lo = make_smem_desc(st + 16384, 16, 1024);
          if (elect_one_sync()) {
            uint32_t accum = 0;
            if (t3) {
              for (int kk = 0; kk < kk1; ++kk) { umma_f16(d_acc, a_lo + ((kk * 32) >> 4), b_hi + ((kk * 32) >> 4), idesc128, accum); accum = 1; }
              for (int kk = 0; kk < kk1; ++kk) umma_f16(d_acc, a_hi + ((kk * 32) >> 4), b_lo + ((kk * 32) >> 4), idesc128, 1u);
            }
            for (int kk = 0; kk < kk1; ++kk) { umma_f16(d_acc, a_hi + ((kk * 32) >> 4), b_hi + ((kk * 32) >> 4), idesc128, accum); accum = 1; }
            umma_commit(w_empty(s));
            umma_commit(tfull);
            if (nt == 1 && j == nck - 1) umma_commit(x2_free);
          }
          __syncwarp();
        }
        // ---- second / third Dense: A = h1 / h2 in tensor memory ----
        for (int i = 0; i < n_seq; ++i, ++it) {
          const int e = seq[i];
          const int s = it % S;
          const uint32_t st = base + C::OFF_W + s * C::STAGE;
          const uint64_t b_hi = make_smem_desc(st, 16, 1024), b_lo = make_smem_desc(st + 16384, 16, 1024);
          if (!(e & 0x80)) {
            const int j = (e >> 2) & 3, k = e & 3;
            // K slab k of h1 is ready as soon as the first-Dense epilogue has written that 64-column half
            mbar_wait3(tempty, (sl & 1) ^ 1, w_full(s), (it / S) & 1, j == 0 ? h1_ready(k) : 0u, ncount & 1);
            tc_fence_after();
            const uint32_t a_hi = tmem_base + C::TM_H1HI + k * 32, a_lo = tmem_base + C::TM_H1LO + k * 32;
            if (elect_one_sync()) {
              if (t3) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_f16_ts(d_acc, a_lo + kk * 8, b_hi + ((kk * 32) >> 4), idesc128, kk > 0 ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_f16_ts(d_acc, a_hi + kk * 8, b_lo + ((kk * 32) >> 4), idesc128, 1u);
              }
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16_ts(d_acc, a_hi + kk * 8, b_hi + ((kk * 32) >> 4), idesc128, (t3 || kk > 0) ? 1u : 0u);
              umma_commit(w_empty(s));
              umma_commit(tfull);
            }
            __syncwarp();
            ++sl;
          } else {
            const uint32_t acc = sl3 & 1;
            mbar_wait3(tempty3(acc), ((sl3 >> 1) & 1) ^ 1, w_full(s), (it / S) & 1, h2_ready, cc & 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + C::TM_ACC3 + acc * 32;
            const uint32_t a_hi = tmem_base + C::TM_H2HI, a_lo = tmem_base + C::TM_H2LO;
            if (elect_one_sync()) {
              if (t3) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_f16_ts(d_tmem, a_lo + kk * 8, b_hi + ((kk * 32) >> 4), idesc32, kk > 0 ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_f16_ts(d_tmem, a_hi + kk * 8, b_lo + ((kk * 32) >> 4), idesc32, 1u);
              }
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16_ts(d_tmem, a_hi + kk * 8, b_hi + ((kk * 32) >> 4), idesc32, (t3 || kk > 0) ? 1u : 0u);
              umma_commit(w_empty(s));
              umma_commit(h2_free);
              umma_commit(tfull3(acc));
            }
            __syncwarp();
            ++sl3; ++cc;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue warps ===========================
    const int t = threadIdx.x - C::EPI0;
    const int quarter = warp & 3;
    const int g = (warp - 4) >> 2;                 // owns columns [16 g, 16 g + 16) of EACH 64-column half of a 128-column chunk
    const int rloc = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int d = p.d, dh = p.d >> 1;
    const int h_ld = p.h_ld;
    const int bits_ld = h_ld >> 4;
    uint8_t* stg_warp = smem_raw + C::OFF_STG + (warp - 4) * 2048;
    uint16_t* s_bits = reinterpret_cast<uint16_t*>(smem_raw + C::OFF_BITS);
    const float* Xs = reinterpret_cast<const float*>(smem_raw + C::OFF_X);
    float run_max = 0.f;
    uint32_t tcount = 0, sl = 0, sl3 = 0, cc = 0;
    // element -> (row, column pair) of the coalesced passes over a tile: thread t handles idx = t + 512 i; when 512 is a
    // multiple of d/2 the column pair is loop invariant and the row advances by a constant (no divisions in the loops)
    const bool reg_idx = (512 % dh) == 0;
    const int jp0 = t % dh, r0i = t / dh, rstep = 512 / dh;
    int ev = 0;
#define F2_EV(tag) do { if (p.dbg && blockIdx.x == 0 && t == 0 && ev < 511) { p.dbg[512 + ev] = clock64(); p.dbg[1024 + ev] = (tag); ++ev; } } while (0)
    auto scatter_x2 = [&](int64_t tile_s, uint32_t tc) {
      const int64_t r0 = tile_s * 128;
      const int rows_s = (int)((p.n - r0) < 128 ? (p.n - r0) : 128);
      mbar_wait(x_full, tc & 1);
      F2_EV(611);
      if (tc > 0) {
        mbar_wait(x2_free, (tc - 1) & 1);
        F2_EV(612);
        if (t == 0) tma_store_wait_read();
        F2_EV(613);
        epi_bar_sync(1, 512);
        F2_EV(614);
      }
      const int n_it = (128 * dh + 511) / 512;
      for (int i = 0; i < n_it; ++i) {
        const int idx = t + 512 * i;
        if (idx >= 128 * dh) break;
        const int r = reg_idx ? r0i + i * rstep : idx / dh, jp = reg_idx ? jp0 : idx - (idx / dh) * dh;
        const float2 x = (r < rows_s) ? *reinterpret_cast<const float2*>(Xs + r * d + 2 * jp) : make_float2(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int k = s_pos2[2 * jp + u];
          if (k >= 0) {
            const float v = (u ? x.y : x.x) * s_x2;
            const __half h = __float2half_rn(v);
            const __half l = __float2half_rn(v - __half2float(h));
            const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
            *reinterpret_cast<__half*>(smem_raw + C::OFF_X2 + off) = h;
            *reinterpret_cast<__half*>(smem_raw + C::OFF_X2 + C::X2_PLANE + off) = l;
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(x2_ready); mbar_arrive(x_empty); }     // the X buffer is free again: next tile's rows may land
      F2_EV(615);
      epi_bar_sync(1, 512);
      F2_EV(616);
      if (t == 0) {
        tma_store_3d(&maps.x2, base + C::OFF_X2, 0, (int)r0, 0);
        if (p.terms > 1) tma_store_3d(&maps.x2, base + C::OFF_X2 + C::X2_PLANE, 0, (int)r0, 1);
      }
    };
    if ((int64_t)blockIdx.x < num_tiles) scatter_x2(blockIdx.x, 0);
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
      const int64_t row0 = tile * 128;
      const int rows_here = (int)((p.n - row0) < 128 ? (p.n - row0) : 128);

      for (int nt = 0; nt < 2; ++nt) {
        const FusedNet& N = p.net[nt];
        const float* b1 = s_bias + nt * 544, * b2 = b1 + 256, * b3 = b1 + 512;
        const float bound1 = amax_x * N.w_sc[0][1] + N.w_sc[0][3];
        const float s_h1 = pow2_scale(bound1 * 1.001f);
        const float bound2 = bound1 * N.w_sc[1][1] + N.w_sc[1][3];
        const float s_h2 = pow2_scale(bound2 * 1.001f);
        const float d1 = 1.f / (s_x2 * N.w_sc[0][0]), d2 = 1.f / (s_h1 * N.w_sc[1][0]), d3 = 1.f / (s_h2 * N.w_sc[2][0]);
        const float ds1 = fmaf(d1, p.rz[0], d1) * s_h1, ds2 = fmaf(d2, p.rz[1], d2) * s_h2, ds3 = fmaf(d3, p.rz[2], d3);
        if (blockIdx.x == 0 && t == 0 && tcount == 0) {
          N.h_meta[0][0] = s_h1; N.h_meta[0][1] = bound1;
          N.h_meta[1][0] = s_h2; N.h_meta[1][1] = bound2;
          if (nt == 0) { p.x2_meta[0] = s_x2; p.x2_meta[1] = amax_x; }
        }
        if (t == 0) tma_store_wait_read();           // the sign-bit staging of the previous network has been read
        epi_bar_sync(2, 512);
        // ---- first Dense epilogue: 128-column chunk j of h1 -> TMEM operand planes + stash ----
        for (int j = 0; j < nck; ++j, ++sl) {
          F2_EV(100 + j);
          if (lane == 0) mbar_wait_relaxed(tfull, sl & 1);
          __syncwarp();
          F2_EV(110 + j);
          tc_fence_after();
          uint32_t va[16], vb[16];
          tmem_ld16x2(tmem_base + lane_off + C::TM_ACC + g * 16, tmem_base + lane_off + C::TM_ACC + 64 + g * 16, va, vb);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty);
          F2_EV(120 + j);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int col = j * 128 + hf * 64 + g * 16;
            if (col >= h_ld) break;                  // warp-uniform: a 64- or 192-wide hidden layer has no second half here
            float acc[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] = __uint_as_float(hf ? vb[q] : va[q]);
            uint32_t hi[8], lo[8];
            const uint32_t bits = f2_act_split(acc, ds1, b1 + col, hi, lo);
            tmem_st8(tmem_base + lane_off + C::TM_H1HI + (col >> 1), hi);
            tmem_st8(tmem_base + lane_off + C::TM_H1LO + (col >> 1), lo);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(h1_ready(2 * j + hf));     // the second Dense may consume this K slab now
            F2_EV(130 + hf);
            if (!(p.dbg_flags & 1)) {
              fused_stash_store(stg_warp, &maps.h[nt][0], hi, lo, lane, col, (int)row0 + quarter * 32, p.terms > 1);
              s_bits[rloc * bits_ld + (col >> 4)] = (uint16_t)bits;
            }
          }
          F2_EV(150 + j);
        }
        // ---- second Dense (slabs summed in registers) and third Dense ----
        float racc[2][16], racc3[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) racc3[q] = 0.f;
        for (int i = 0; i < n_seq; ++i) {
          const int e = seq[i];
          if (!(e & 0x80)) {
            const int j = (e >> 2) & 3, k = e & 3;
            F2_EV(200 + 4 * j + k);
            if (lane == 0) mbar_wait_relaxed(tfull, sl & 1);
            __syncwarp();
            F2_EV(220 + 4 * j + k);
            tc_fence_after();
            uint32_t va[16], vb[16];
            tmem_ld16x2(tmem_base + lane_off + C::TM_ACC + g * 16, tmem_base + lane_off + C::TM_ACC + 64 + g * 16, va, vb);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
            ++sl;
            if (k == 0) {
#pragma unroll
              for (int q = 0; q < 16; ++q) { racc[0][q] = __uint_as_float(va[q]); racc[1][q] = __uint_as_float(vb[q]); }
            } else {
#pragma unroll
              for (int q = 0; q < 16; ++q) { racc[0][q] += __uint_as_float(va[q]); racc[1][q] += __uint_as_float(vb[q]); }
            }
            if (k == nks - 1) {
              // both 64-column halves of the chunk, one after the other through the single operand buffer: every warp works
              // on half a first, so the third-Dense MMAs of half a run while half b is still being activated and split
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                const int col = j * 128 + hf * 64 + g * 16;
                if (col >= h_ld) break;
                uint32_t hi[8], lo[8];
                const uint32_t bits = f2_act_split(racc[hf], ds2, b2 + col, hi, lo);
                F2_EV(300 + hf);
                if (cc > 0 && lane == 0) mbar_wait(h2_free, (cc - 1) & 1);
                __syncwarp();
                F2_EV(310 + hf);
                tc_fence_after();
                tmem_st8(tmem_base + lane_off + C::TM_H2HI + g * 8, hi);
                tmem_st8(tmem_base + lane_off + C::TM_H2LO + g * 8, lo);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(h2_ready);
                ++cc;
                F2_EV(320 + hf);
                if (!(p.dbg_flags & 1)) {
                  fused_stash_store(stg_warp, &maps.h[nt][1], hi, lo, lane, col, (int)row0 + quarter * 32, p.terms > 1);
                  s_bits[128 * bits_ld + rloc * bits_ld + (col >> 4)] = (uint16_t)bits;
                }
              }
            }
          } else {
            const uint32_t acc = sl3 & 1;
            F2_EV(400);
            if (lane == 0) mbar_wait_relaxed(tfull3(acc), (sl3 >> 1) & 1);
            __syncwarp();
            F2_EV(410);
            tc_fence_after();
            uint32_t v[8];
            tmem_ld8(tmem_base + lane_off + C::TM_ACC3 + acc * 32 + g * 8, v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty3(acc));
            ++sl3;
#pragma unroll
            for (int q = 0; q < 8; ++q) racc3[q] += __uint_as_float(v[q]);
          }
        }
        F2_EV(500);
        if (lane == 0) tma_store_wait_read();
        fence_proxy_async();
        epi_bar_sync(2, 512);
        F2_EV(510);
        if (t == 0 && !(p.dbg_flags & 1)) {
          const uint32_t bytes = 128u * (uint32_t)bits_ld * 2u;
          for (int l = 0; l < 2; ++l) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(N.h_bits[l] + row0 * bits_ld), "r"(base + C::OFF_BITS + l * 128 * bits_ld * 2), "r"(bytes) : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        {
          float part = 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int k = g * 8 + q;
            float a = fmaf(racc3[q], ds3, b3[k]);
            if (nt == 0) { a = tanhf(a); if (k < p.c) part += a; s_S[rloc * C::ST_LD + k] = a; }
            else s_T[rloc * C::ST_LD + k] = a;
          }
          if (nt == 0) atomicAdd(&s_ld[rloc], part);
        }
      }
      F2_EV(600);
      epi_bar_sync(1, 512);
      F2_EV(610);
      if (tile + gridDim.x < num_tiles) scatter_x2(tile + gridDim.x, tcount + 1);
      F2_EV(620);
      // ---- coupling arithmetic: coalesced pass; the tile's rows are read again from global memory (L2 resident) ----
      {
        // all loads first (one round trip to L2 instead of one per iteration), then the arithmetic
        float2 xv[8];
        float ldv = 0.f;
        if (t < rows_here && p.ld) ldv = p.ld[row0 + t];
        const float* xrow0 = p.Xin + row0 * d;
        float* yrow0 = p.Xout + row0 * d;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = t + 512 * i;
          const int r = reg_idx ? r0i + i * rstep : idx / dh, jp = reg_idx ? jp0 : idx - (idx / dh) * dh;
          xv[i] = (idx < 128 * dh && r < rows_here) ? *reinterpret_cast<const float2*>(xrow0 + r * d + 2 * jp) : make_float2(0.f, 0.f);
        }
        F2_EV(621);
        if (xv[0].x == 12345.678f) run_max = 1.f;      // (consume the loads here so that the event above marks their arrival)
        F2_EV(622);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = t + 512 * i;
          const int r = reg_idx ? r0i + i * rstep : idx / dh, jp = reg_idx ? jp0 : idx - (idx / dh) * dh;
          if (idx >= 128 * dh || r >= rows_here) continue;
          float y[2] = {xv[i].x, xv[i].y};
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int k = s_pos[2 * jp + u];
            if (k >= 0) {
              const float sv = s_S[r * C::ST_LD + k], tv = s_T[r * C::ST_LD + k];
              y[u] = p.inv ? (y[u] - tv) * expf(-sv) : fmaf(expf(sv), y[u], tv);
            }
            run_max = fmaxf(run_max, fabsf(y[u]));
          }
          *reinterpret_cast<float2*>(yrow0 + r * d + 2 * jp) = make_float2(y[0], y[1]);
        }
        F2_EV(623);
        if (t < 128) {
          if (t < rows_here && p.ld) p.ld[row0 + t] = ldv + (p.inv ? -s_ld[t] : s_ld[t]);
          s_ld[t] = 0.f;
        }
      }
      if (p.net[0].out) {
        const int c = p.c;
        if ((512 % c) == 0) {
          const int k = t % c, rs = 512 / c;
          float* o = p.net[0].out + row0 * c;
          for (int r = t / c; r < rows_here; r += rs) o[r * c + k] = s_S[r * C::ST_LD + k];
        } else {
          for (int idx = t; idx < rows_here * c; idx += 512) {
            const int r = idx / c, k = idx - r * c;
            p.net[0].out[row0 * c + idx] = s_S[r * C::ST_LD + k];
          }
        }
      }
      if (p.net[1].out)
        for (int idx = t; idx < rows_here * p.c; idx += 512) {
          const int r = idx / p.c, k = idx - r * p.c;
          p.net[1].out[row0 * p.c + idx] = s_T[r * C::ST_LD + k];
        }
      F2_EV(630);
    }
    if (p.y_meta) {
      run_max = warp_max(run_max);
      if (lane == 0) meta_amax(p.y_meta, run_max);
    }
    // every thread that issued bulk stores (stash pieces, sign bits) waits for them before the CTA's shared memory goes away
    __syncwarp();
    if (elect_one_sync()) tma_store_wait_all();
    if (t == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
