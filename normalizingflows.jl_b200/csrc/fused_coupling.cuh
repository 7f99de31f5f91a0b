// K3f -- fused conditioner of an AffineCoupling (reference src/flows/realnvp.jl:57-110, src/flows/utils.jl:71-100).
//
// One persistent CTA per SM carries a 128-sample tile through BOTH conditioner networks (s and t), all three Dense
// layers each, and the coupling arithmetic, without a hidden activation ever visiting HBM as an operand:
//
//   X tile (fp32, one 1-D bulk copy) -> x2 = X[:, idx2] split into fp16 hi/lo planes (smem, UMMA K-major)
//   L1: acc[128 x 64] = x2 W1[chunk]^T        A from smem,  B (weights) streamed by TMA through a 16 KB-stage ring
//       epilogue: bias + leakyrelu + hi/lo split -> TMEM (tcgen05.st) as the A operand of L2   (+ stash, see below)
//   L2: acc[128 x 64] = h1 W2[chunk, slab]^T  A from TMEM (tcgen05.mma .ts form), one K slab of 64 per accumulator,
//       slabs summed in registers with round-to-nearest adds (the tensor core truncates its fp32 accumulator)
//       epilogue: bias + leakyrelu + split -> TMEM chunk buffer = A operand of L3
//   L3: acc[128 x 32] += h2[chunk] W3[:, chunk]^T, drained per slab; s = tanh(.), t = (.)
//   coupling: y1 = exp(s) x1 + t, logdet += sum(s)   (inverse direction: x1 = (y1 - t) exp(-s), logdet -= sum(s))
//
// TMEM (512 columns): [0,128) h1 hi | [128,256) h1 lo | [256,384) two 64-column accumulators |
//                     [384,416) h2 chunk hi | [416,448) h2 chunk lo | [448,512) two 32-column L3 accumulators
// fp16 operands in TMEM are packed two per 32-bit column (element k of row m: lane m, column k/2, half k%2).
//
// Warp roles: warp 0 producer (TMA weights, bulk copy of X tiles), warp 1 MMA issuer (one elected thread), warps 2-3 idle,
// warps 4..19 epilogue (TMEM lane quarter = warp % 4, 16-column group = (warp - 4) / 4).
//
// What still goes to HBM is the stash the backward pass consumes (same formats as the layer-by-layer path, so
// tc_mlp_backward works unchanged): x2 planes, h1 / h2 planes + sign bits per network, s (fp32), and the new state.
//
// This file is included by tc_gemm.cu inside namespace nf { namespace { ... } }.

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Wait for up to three mbarrier phases at once: the try_waits are independent, so their ~100-200 cycle latencies overlap
// (three back-to-back mbar_wait calls, or three lanes polling one barrier each, serialise instead).  bar == 0: skip.
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done;
}
__device__ __forceinline__ void mbar_wait3(uint32_t b0, uint32_t p0, uint32_t b1, uint32_t p1, uint32_t b2, uint32_t p2) {
  uint32_t d0 = 0, d1 = 0, d2 = b2 ? 0u : 1u, spins = 0;
  while (true) {
    const uint32_t t0 = d0 ? 1u : mbar_try(b0, p0);
    const uint32_t t1 = d1 ? 1u : mbar_try(b1, p1);
    const uint32_t t2 = d2 ? 1u : mbar_try(b2, p2);
    d0 = t0; d1 = t1; d2 = t2;
    if (d0 & d1 & d2) break;
    if (++spins > (1u << 26)) __trap();
  }
}
// one lane of a converged warp (cute::elect_one_sync)
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, %2;\n\t"
      "@%%px mov.s32 %1, 1;\n\t"
      "mov.s32 %0, %%rx;\n\t}"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void epi_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct FusedNet {
  const float* w_sc[3];     // per-Dense scalars {weight scale, max row L1, max col L1, max |b|}
  const float* bias[3];     // zero padded fp32
  __half* h_planes[2];      // stash: hidden activations of Dense 1 / 2 (hi plane; lo at + h_plane_elems)
  uint16_t* h_bits[2];      // packed (pre-activation > 0) bits, 16 columns per word, row stride h_ld / 16 words
  float* h_meta[2];         // {scale, bound} slots the backward kernels read
  float* out;               // fp32 [n, c]: s after tanh / t (may be null for t)
};

struct FusedFwdParams {
  int64_t n;
  int d, c, cbar;
  int nch;                  // hidden width / 64 (padded), 1..4
  int kk1;                  // K = 16 steps of the first Dense that carry data
  int inv;
  int terms;                // 1 or 3
  const float* Xin;
  float* Xout;
  float* ld;
  const int* pos;           // [d] position inside idx1 or -1
  const int* pos2;          // [d] position inside idx2 or -1
  const float* x_meta;      // {., exact max |Xin|}
  float* y_meta;            // exact max |Xout| accumulates here (may be null)
  float* x2_meta;           // {scale, amax} of the stashed x2 planes
  int64_t h_plane_elems;    // rows_pad * h_ld
  int h_ld;                 // 64 * nch
  float rz[3];
  int dbg_flags;            // experiments (NFCUDA_DBG_FLAGS): 1 skip the hidden-activation stash stores
  long long* dbg;           // optional clock64 timeline of CTA 0: [role 0..2][512]
  FusedNet net[2];
};

#define NF_FDBG(role, idx) do { if (p.dbg && blockIdx.x == 0 && (idx) < 512) p.dbg[(role) * 512 + (idx)] = clock64(); } while (0)

struct FusedFwdMaps {
  CUtensorMap w[2][3];      // weight planes, K-major, box {64, 64 | 32, 1}
  CUtensorMap x2;           // stash of the x2 planes: box {64, 128, 1}
  CUtensorMap h[2][2];      // stash of the hidden planes [net][layer]: box {16, 32, 1}, SWIZZLE_32B
  CUtensorMap w128[2][2];   // wide variant: first / second Dense weight planes with box {64, 128, 1}
};

struct FusedCfg {
  static constexpr int STAGES = 4;
  static constexpr int STAGE = 16384;            // weight slab: 64 rows x 128 B, hi plane then lo plane (+8192)
  static constexpr int XS = 32768;               // one X tile: 128 rows x d floats (d <= 64)
  static constexpr int X2_PLANE = 16384;         // 128 rows x 128 B
  static constexpr int ST_LD = 33;               // padded row stride of the s / t staging tiles (floats)
  static constexpr int OFF_W = 0;
  static constexpr int OFF_X = OFF_W + STAGES * STAGE;
  static constexpr int OFF_X2 = OFF_X + 2 * XS;
  static constexpr int OFF_S = OFF_X2 + 2 * X2_PLANE;
  // per-warp staging of the hidden-activation stash: 16 warps x {hi, lo} x [32 rows x 32 B] (SWIZZLE_32B), left by TMA
  // stores; the t staging tile aliases it (written once per tile, after every warp's stash stores have read their tiles)
  static constexpr int OFF_STG = OFF_S + 128 * ST_LD * 4;
  static constexpr int OFF_T = OFF_STG;
  static constexpr int OFF_BITS = OFF_STG + 16 * 2048;            // [2 layers][128 rows][8 * nch B] sign bits of one network
  static constexpr int OFF_BIAS = OFF_BITS + 2 * 4096;            // [2 nets][256 + 256 + 32] floats
  static constexpr int OFF_LD = OFF_BIAS + 2 * 544 * 4;           // [128] floats
  static constexpr int OFF_BAR = OFF_LD + 512;
  static constexpr int N_BARS = 2 * STAGES + 4 /*x full/empty*/ + 2 /*x2 ready/free*/ + 4 /*tfull/tempty*/ + 4 /*tfull3/tempty3*/ +
                                4 /*h1 ready*/ + 2 /*h2 ready/free*/;
  static constexpr int OFF_SEQ = OFF_BAR + 8 * N_BARS + 16;
  static constexpr int SMEM = OFF_SEQ + 64;
  static constexpr int EPI0 = 128;                // first epilogue thread: warp 0 producer, 1 MMA issuer, 2-3 idle, 4..19 epilogue
  static constexpr int THREADS = EPI0 + 512;
  // TMEM columns
  static constexpr int TM_H1HI = 0, TM_H1LO = 128, TM_ACC = 256, TM_H2HI = 384, TM_H2LO = 416, TM_ACC3 = 448;
};

// item of the per-network MMA schedule after the first Dense: L2 slab (j, k) or L3 slab j
__device__ __forceinline__ int fused_build_seq(uint8_t* seq, int nch) {
  int n = 0;
  const int kk = nch > 1 ? 1 : 0;
  for (int j = 0; j < nch; ++j)
    for (int k = 0; k < nch; ++k) {
      seq[n++] = (uint8_t)((j << 2) | k);
      if (j > 0 && k == kk) seq[n++] = (uint8_t)(0x80 | (j - 1));
    }
  seq[n++] = (uint8_t)(0x80 | (nch - 1));
  return n;
}


// One warp leaves its [32 rows x 16 columns] piece of a hidden-activation chunk: registers -> the warp's staging tile
// (SWIZZLE_32B pattern, conflict-free 16-byte stores) -> one TMA store per plane.  Rows past the batch are clipped by the map.
__device__ __forceinline__ void fused_stash_store(uint8_t* stg, const CUtensorMap* map, const uint32_t (&hi)[8], const uint32_t (&lo)[8],
                                                  int lane, int col0, int row_base, bool two_planes) {
  // one ELECTED lane waits / issues (elect.sync keeps the bulk-tensor instructions on the uniform datapath; a `lane == 0`
  // branch costs an election loop per instruction)
  if (elect_one_sync()) tma_store_wait_read();   // the previous stores of this warp have read the tile
  __syncwarp();
  const int sw = (lane >> 2) & 1;
  uint8_t* rowp = stg + lane * 32;
  *reinterpret_cast<uint4*>(rowp + ((0 ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(rowp + ((1 ^ sw) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  if (two_planes) {
    *reinterpret_cast<uint4*>(rowp + 1024 + ((0 ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(rowp + 1024 + ((1 ^ sw) << 4)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
  }
  fence_proxy_async();
  __syncwarp();
  if (elect_one_sync()) {
    tma_store_3d(map, smem_u32(stg), col0, row_base, 0);
    if (two_planes) tma_store_3d(map, smem_u32(stg) + 1024, col0, row_base, 1);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(FusedCfg::THREADS, 1)
fused_affine_fwd_kernel(const __grid_constant__ FusedFwdMaps maps, const FusedFwdParams p) {
  using C = FusedCfg;
  constexpr int S = C::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  const uint32_t bars = base + C::OFF_BAR;
  auto w_full = [&](int s) { return bars + 8u * s; };
  auto w_empty = [&](int s) { return bars + 8u * (S + s); };
  auto x_full = [&](int s) { return bars + 8u * (2 * S + s); };
  auto x_empty = [&](int s) { return bars + 8u * (2 * S + 2 + s); };
  const uint32_t x2_ready = bars + 8u * (2 * S + 4), x2_free = bars + 8u * (2 * S + 5);
  auto tfull = [&](int a) { return bars + 8u * (2 * S + 6 + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * S + 8 + a); };
  auto tfull3 = [&](int a) { return bars + 8u * (2 * S + 10 + a); };
  auto tempty3 = [&](int a) { return bars + 8u * (2 * S + 12 + a); };
  auto h1_ready = [&](int k) { return bars + 8u * (2 * S + 14 + k); };
  const uint32_t h2_ready = bars + 8u * (2 * S + 18), h2_free = bars + 8u * (2 * S + 19);
  const uint32_t tmem_slot = bars + 8u * C::N_BARS;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + C::OFF_BAR + 8 * C::N_BARS);
  uint8_t* seq = smem_raw + C::OFF_SEQ;
  float* s_bias = reinterpret_cast<float*>(smem_raw + C::OFF_BIAS);
  float* s_ld = reinterpret_cast<float*>(smem_raw + C::OFF_LD);
  float* s_S = reinterpret_cast<float*>(smem_raw + C::OFF_S);
  float* s_T = reinterpret_cast<float*>(smem_raw + C::OFF_T);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int64_t num_tiles = (p.n + 127) / 128;
  const int nch = p.nch;
  const int n_seq = nch * nch + nch;

  // ---- scales (every thread derives the same values from the same device scalars) ----
  const float amax_x = __uint_as_float(reinterpret_cast<const unsigned int*>(p.x_meta)[1]);
  const float s_x2 = pow2_scale(amax_x);

  if (warp == 0 && lane == 0) {
    for (int nt = 0; nt < 2; ++nt)
      for (int l = 0; l < 3; ++l) tma_prefetch_desc(&maps.w[nt][l]);
    tma_prefetch_desc(&maps.x2);
    for (int s = 0; s < S; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(x_full(s), 1); mbar_init(x_empty(s), 16); }
    mbar_init(x2_ready, 16); mbar_init(x2_free, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 16); mbar_init(tfull3(a), 1); mbar_init(tempty3(a), 16); }
    for (int k = 0; k < 4; ++k) mbar_init(h1_ready(k), 16);
    mbar_init(h2_ready, 16); mbar_init(h2_free, 1);
    fence_barrier_init();
    fused_build_seq(seq, nch);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (threadIdx.x >= C::EPI0) {
    const int t = threadIdx.x - C::EPI0;
    // zero the x2 planes once (padding columns stay zero), the logdet staging, and fetch the biases
    for (int i = t; i < 2 * C::X2_PLANE / 16; i += 512) reinterpret_cast<uint4*>(smem_raw + C::OFF_X2)[i] = make_uint4(0, 0, 0, 0);
    if (t < 128) s_ld[t] = 0.f;
    for (int i = t; i < 2 * 544; i += 512) {
      const int nt = i / 544, o = i % 544;
      // hidden-layer biases are stored already multiplied by the scale of the activation they feed (one FFMA per element later)
      const float bound1 = amax_x * p.net[nt].w_sc[0][1] + p.net[nt].w_sc[0][3];
      const float bound2 = bound1 * p.net[nt].w_sc[1][1] + p.net[nt].w_sc[1][3];
      float v;
      if (o < 256) v = o < 64 * nch ? p.net[nt].bias[0][o] * pow2_scale(bound1 * 1.001f) : 0.f;
      else if (o < 512) v = (o - 256) < 64 * nch ? p.net[nt].bias[1][o - 256] * pow2_scale(bound2 * 1.001f) : 0.f;
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
    // warp-uniform like the MMA issuer: every lane walks the schedule, one elected lane arms the barrier and issues the copies
    {
      uint32_t it = 0, xt = 0;
      auto load_x = [&](int64_t tile) {
        const int xs = xt & 1;
        mbar_wait(x_empty(xs), ((xt >> 1) & 1) ^ 1);
        const int64_t r0 = tile * 128;
        const int rows = (int)((p.n - r0) < 128 ? (p.n - r0) : 128);
        const uint32_t bytes = (uint32_t)rows * (uint32_t)p.d * 4u;
        if (elect_one_sync()) {
          mbar_expect_tx(x_full(xs), bytes);
          bulk_load_1d(base + C::OFF_X + xs * C::XS, p.Xin + r0 * p.d, bytes, x_full(xs));
        }
        __syncwarp();
        ++xt;
      };
      auto load_w = [&](const CUtensorMap* map, int c0, int c1, uint32_t bytes_per_plane) {
        const int s = it % S;
        NF_FDBG(2, 4 * it);
        mbar_wait(w_empty(s), ((it / S) & 1) ^ 1);
        NF_FDBG(2, 4 * it + 1);
        const uint32_t st = base + C::OFF_W + s * C::STAGE;
        if (elect_one_sync()) {
          mbar_expect_tx(w_full(s), (p.terms > 1 ? 2u : 1u) * bytes_per_plane);
          tma_load_3d(st, map, w_full(s), c0, c1, 0);
          if (p.terms > 1) tma_load_3d(st + 8192, map, w_full(s), c0, c1, 1);
        }
        __syncwarp();
        NF_FDBG(2, 4 * it + 2);
        ++it;
      };
      if ((int64_t)blockIdx.x < num_tiles) load_x(blockIdx.x);
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int nt = 0; nt < 2; ++nt) {
          for (int j = 0; j < nch; ++j) load_w(&maps.w[nt][0], 0, j * 64, 8192);
          for (int i = 0; i < n_seq; ++i) {
            const int e = seq[i];
            if (e & 0x80) load_w(&maps.w[nt][2], (e & 3) * 64, 0, 4096);
            else load_w(&maps.w[nt][1], (e & 3) * 64, ((e >> 2) & 3) * 64, 8192);
          }
          // the next X tile: by now the epilogue warps have long released the other staging buffer
          if (nt == 0 && tile + gridDim.x < num_tiles) load_x(tile + gridDim.x);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // (a second issuer warp alternating over the items was measured: no gain -- the tensor pipe itself is the pace setter,
    //  a 128 x 64 x 16 MMA with A in tensor memory takes ~65 cycles, twice its nominal 32, because the A tile is re-read per MMA)
    // The whole warp walks the schedule (uniform control flow, barrier waits by every lane) and ONE elected lane issues:
    // with `elect.sync` the compiler keeps descriptors in uniform registers and emits back-to-back UTCHMMA; a plain
    // `lane == 0` branch costs an election loop (~60 cycles) per MMA, more than a 128 x 64 x 16 MMA takes to execute.
    {
      constexpr uint32_t idesc64 = make_idesc(128, 64, 0, 0);
      constexpr uint32_t idesc32 = make_idesc(128, 32, 0, 0);
      uint32_t it = 0, sl = 0, sl3 = 0, tcount = 0, ncount = 0, cc = 0;
      const bool t3 = p.terms > 1;
      const uint32_t x2a = base + C::OFF_X2;
      const int kk1 = p.kk1;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
        mbar_wait(x2_ready, tcount & 1);
        tc_fence_after();
        for (int nt = 0; nt < 2; ++nt, ++ncount) {
          // ---- first Dense: A = x2 planes in shared memory ----
          for (int j = 0; j < nch; ++j, ++it, ++sl) {
            const int s = it % S;
            const uint32_t acc = sl & 1;
            const bool last_l1 = (nt == 1 && j == nch - 1);
            mbar_wait3(tempty(acc), ((sl >> 1) & 1) ^ 1, w_full(s), (it / S) & 1, 0, 0);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + C::TM_ACC + acc * 64;
            const uint32_t st = base + C::OFF_W + s * C::STAGE;
            const uint64_t a_hi = make_smem_desc(x2a, 16, 1024), a_lo = make_smem_desc(x2a + C::X2_PLANE, 16, 1024);
            const uint64_t b_hi = make_smem_desc(st, 16, 1024), b_lo = make_smem_desc(st + 8192, 16, 1024);
            if (elect_one_sync()) {
              uint32_t accum = 0;
              if (t3) {
                for (int kk = 0; kk < kk1; ++kk) { umma_f16(d_tmem, a_lo + ((kk * 32) >> 4), b_hi + ((kk * 32) >> 4), idesc64, accum); accum = 1; }
                for (int kk = 0; kk < kk1; ++kk) umma_f16(d_tmem, a_hi + ((kk * 32) >> 4), b_lo + ((kk * 32) >> 4), idesc64, 1u);
              }
              for (int kk = 0; kk < kk1; ++kk) { umma_f16(d_tmem, a_hi + ((kk * 32) >> 4), b_hi + ((kk * 32) >> 4), idesc64, accum); accum = 1; }
              umma_commit(w_empty(s));
              umma_commit(tfull(acc));
              if (last_l1) umma_commit(x2_free);
            }
            __syncwarp();
          }
          // ---- second / third Dense: A = h1 / h2 in tensor memory ----
          for (int i = 0; i < n_seq; ++i, ++it) {
            const int e = seq[i];
            const int s = it % S;
            const uint32_t st = base + C::OFF_W + s * C::STAGE;
            const uint64_t b_hi = make_smem_desc(st, 16, 1024), b_lo = make_smem_desc(st + 8192, 16, 1024);
            if (!(e & 0x80)) {
              const int j = (e >> 2) & 3, k = e & 3;
              const uint32_t acc = sl & 1;
              NF_FDBG(0, 4 * it);
              // the three conditions are polled by three lanes at once (a completed try_wait still costs ~100 cycles)
              mbar_wait3(tempty(acc), ((sl >> 1) & 1) ^ 1, w_full(s), (it / S) & 1, j == 0 ? h1_ready(k) : 0u, ncount & 1);
              NF_FDBG(0, 4 * it + 1);
              tc_fence_after();
              const uint32_t d_tmem = tmem_base + C::TM_ACC + acc * 64;
              const uint32_t a_hi = tmem_base + C::TM_H1HI + k * 32, a_lo = tmem_base + C::TM_H1LO + k * 32;
              if (elect_one_sync()) {
                if (t3) {
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk) umma_f16_ts(d_tmem, a_lo + kk * 8, b_hi + ((kk * 32) >> 4), idesc64, kk > 0 ? 1u : 0u);
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk) umma_f16_ts(d_tmem, a_hi + kk * 8, b_lo + ((kk * 32) >> 4), idesc64, 1u);
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_f16_ts(d_tmem, a_hi + kk * 8, b_hi + ((kk * 32) >> 4), idesc64, (t3 || kk > 0) ? 1u : 0u);
                umma_commit(w_empty(s));
                umma_commit(tfull(acc));
              }
              __syncwarp();
              NF_FDBG(0, 4 * it + 2);
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
    }
  } else if (warp >= 4) {
    // =========================== epilogue warps ===========================
    const int t = threadIdx.x - C::EPI0;           // 0 .. 511
    const int quarter = warp & 3;
    const int g = (warp - 4) >> 2;                 // 16-column group of a 64-column chunk (8-column group of the 32 outputs)
    const int rloc = quarter * 32 + lane;          // row of the tile this thread owns in TMEM
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int d = p.d, dh = p.d >> 1;
    const int h_ld = p.h_ld;
    const int bits_ld = h_ld >> 4;                 // 16-bit words per row of the sign-bit staging (and of the global layout)
    uint8_t* stg_warp = smem_raw + C::OFF_STG + (warp - 4) * 2048;
    uint16_t* s_bits = reinterpret_cast<uint16_t*>(smem_raw + C::OFF_BITS);
    float run_max = 0.f;
    uint32_t tcount = 0, sl = 0, sl3 = 0, cc = 0;
    // x2 = X[:, idx2] * s_x2 -> hi / lo planes (UMMA K-major, SWIZZLE_128B) for tile number `tc` of this CTA, then the stash copy
    auto scatter_x2 = [&](int64_t tile_s, uint32_t tc) {
      const int xs_s = tc & 1;
      const float* Xq = reinterpret_cast<const float*>(smem_raw + C::OFF_X + xs_s * C::XS);
      const int64_t r0 = tile_s * 128;
      const int rows_s = (int)((p.n - r0) < 128 ? (p.n - r0) : 128);
      mbar_wait(x_full(xs_s), (tc >> 1) & 1);
      if (tc > 0) {
        mbar_wait(x2_free, (tc - 1) & 1);          // the previous tile's first-Dense MMAs are complete
        if (t == 0) tma_store_wait_read();         // ... and its x2 stash store has read the planes
        epi_bar_sync(1, 512);
      }
      for (int idx = t; idx < 128 * dh; idx += 512) {
        const int r = idx / dh, jp = idx - r * dh;
        const float2 x = (r < rows_s) ? *reinterpret_cast<const float2*>(Xq + r * d + 2 * jp) : make_float2(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int k = p.pos2[2 * jp + u];
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
      if (lane == 0) mbar_arrive(x2_ready);
      epi_bar_sync(1, 512);
      if (t == 0) {                                // stash of the x2 planes for the weight-gradient kernel
        tma_store_3d(&maps.x2, base + C::OFF_X2, 0, (int)r0, 0);
        if (p.terms > 1) tma_store_3d(&maps.x2, base + C::OFF_X2 + C::X2_PLANE, 0, (int)r0, 1);
      }
    };
    if ((int64_t)blockIdx.x < num_tiles) scatter_x2(blockIdx.x, 0);
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
      const int xs = tcount & 1;
      const float* Xs = reinterpret_cast<const float*>(smem_raw + C::OFF_X + xs * C::XS);
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
        // the sign-bit staging of the previous network has been read by its bulk stores
        if (t == 0) tma_store_wait_read();
        epi_bar_sync(2, 512);
        // ---- first Dense epilogue: chunk j of h1 -> TMEM operand planes + stash ----
        for (int j = 0; j < nch; ++j, ++sl) {
          const uint32_t acc = sl & 1;
          if (lane == 0 || (p.dbg_flags & 4)) mbar_wait_relaxed(tfull(acc), (sl >> 1) & 1);
          __syncwarp();
          tc_fence_after();
          uint32_t v[16];
          tmem_ld16(tmem_base + lane_off + C::TM_ACC + acc * 64 + g * 16, v);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty(acc));
          const int col = j * 64 + g * 16;
          uint32_t hi[8], lo[8], bits = 0;
#pragma unroll
          for (int q = 0; q < 16; q += 2) {
            float a = fmaf(__uint_as_float(v[q]), ds1, b1[col + q]);
            float b = fmaf(__uint_as_float(v[q + 1]), ds1, b1[col + q + 1]);
            bits |= (a > 0.f ? 1u : 0u) << q;
            bits |= (b > 0.f ? 1u : 0u) << (q + 1);
            a = fmaxf(a, 0.01f * a); b = fmaxf(b, 0.01f * b);
            split_pair(a, b, hi[q >> 1], lo[q >> 1]);
          }
          tmem_st8(tmem_base + lane_off + C::TM_H1HI + j * 32 + g * 8, hi);
          tmem_st8(tmem_base + lane_off + C::TM_H1LO + j * 32 + g * 8, lo);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(h1_ready(j));
          if (!(p.dbg_flags & 1)) {
            fused_stash_store(stg_warp, &maps.h[nt][0], hi, lo, lane, col, (int)row0 + quarter * 32, p.terms > 1);
            s_bits[rloc * bits_ld + (col >> 4)] = (uint16_t)bits;
          }
        }
        // ---- second Dense (slabs summed in registers) and third Dense ----
        float racc[16], racc3[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) racc3[q] = 0.f;
        for (int i = 0; i < n_seq; ++i) {
          const int e = seq[i];
          if (!(e & 0x80)) {
            const int j = (e >> 2) & 3, k = e & 3;
            const uint32_t acc = sl & 1;
            if (t == 0) NF_FDBG(1, 4 * sl);
            if (lane == 0 || (p.dbg_flags & 4)) mbar_wait_relaxed(tfull(acc), (sl >> 1) & 1);
            __syncwarp();
            if (t == 0) NF_FDBG(1, 4 * sl + 1);
            tc_fence_after();
            uint32_t v[16];
            tmem_ld16(tmem_base + lane_off + C::TM_ACC + acc * 64 + g * 16, v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(acc));
            if (t == 0) NF_FDBG(1, 4 * sl + 2);
            ++sl;
            if (k == 0) {
#pragma unroll
              for (int q = 0; q < 16; ++q) racc[q] = __uint_as_float(v[q]);
            } else {
#pragma unroll
              for (int q = 0; q < 16; ++q) racc[q] += __uint_as_float(v[q]);
            }
            if (k == nch - 1) {
              const int col = j * 64 + g * 16;
              uint32_t hi[8], lo[8], bits = 0;
#pragma unroll
              for (int q = 0; q < 16; q += 2) {
                float a = fmaf(racc[q], ds2, b2[col + q]);
                float b = fmaf(racc[q + 1], ds2, b2[col + q + 1]);
                bits |= (a > 0.f ? 1u : 0u) << q;
                bits |= (b > 0.f ? 1u : 0u) << (q + 1);
                a = fmaxf(a, 0.01f * a); b = fmaxf(b, 0.01f * b);
                split_pair(a, b, hi[q >> 1], lo[q >> 1]);
              }
              if (cc > 0 && lane == 0) mbar_wait(h2_free, (cc - 1) & 1);
              __syncwarp();    // the third-Dense MMAs of the previous chunk have read the buffer
              tc_fence_after();
              tmem_st8(tmem_base + lane_off + C::TM_H2HI + g * 8, hi);
              tmem_st8(tmem_base + lane_off + C::TM_H2LO + g * 8, lo);
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(h2_ready);
              ++cc;
              if (t == 0) NF_FDBG(1, 4 * (sl - 1) + 3);
              if (!(p.dbg_flags & 1)) {
                fused_stash_store(stg_warp, &maps.h[nt][1], hi, lo, lane, col, (int)row0 + quarter * 32, p.terms > 1);
                s_bits[128 * bits_ld + rloc * bits_ld + (col >> 4)] = (uint16_t)bits;
              }
            }
          } else {
            const uint32_t acc = sl3 & 1;
            if (lane == 0 || (p.dbg_flags & 4)) mbar_wait_relaxed(tfull3(acc), (sl3 >> 1) & 1);
            __syncwarp();
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
        // every warp's stash stores have read their staging tiles (the t tile aliases them) and the sign bits are complete
        if (lane == 0) tma_store_wait_read();
        fence_proxy_async();
        epi_bar_sync(2, 512);
        if (t == 0 && !(p.dbg_flags & 1)) {
          const uint32_t bytes = 128u * (uint32_t)bits_ld * 2u;
          for (int l = 0; l < 2; ++l) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(N.h_bits[l] + row0 * bits_ld), "r"(base + C::OFF_BITS + l * 128 * bits_ld * 2), "r"(bytes) : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        // ---- outputs of this network -> staging tile ----
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
      epi_bar_sync(1, 512);
      // the next tile's conditioner input first: its first-Dense MMAs then run while this tile's coupling arithmetic is done
      const bool early_scatter = !(p.dbg_flags & 2);
      if (early_scatter && tile + gridDim.x < num_tiles) scatter_x2(tile + gridDim.x, tcount + 1);
      // ---- coupling arithmetic: coalesced pass over the X tile ----
      for (int idx = t; idx < 128 * dh; idx += 512) {
        const int r = idx / dh, jp = idx - r * dh;
        if (r >= rows_here) break;
        const float2 x = *reinterpret_cast<const float2*>(Xs + r * d + 2 * jp);
        float y[2] = {x.x, x.y};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int k = p.pos[2 * jp + u];
          if (k >= 0) {
            const float sv = s_S[r * C::ST_LD + k], tv = s_T[r * C::ST_LD + k];
            y[u] = p.inv ? (y[u] - tv) * expf(-sv) : fmaf(expf(sv), y[u], tv);
          }
          run_max = fmaxf(run_max, fabsf(y[u]));
        }
        *reinterpret_cast<float2*>(p.Xout + (row0 + r) * d + 2 * jp) = make_float2(y[0], y[1]);
      }
      if (p.net[0].out)
        for (int idx = t; idx < rows_here * p.c; idx += 512) {
          const int r = idx / p.c, k = idx - r * p.c;
          p.net[0].out[row0 * p.c + idx] = s_S[r * C::ST_LD + k];
        }
      if (p.net[1].out)
        for (int idx = t; idx < rows_here * p.c; idx += 512) {
          const int r = idx / p.c, k = idx - r * p.c;
          p.net[1].out[row0 * p.c + idx] = s_T[r * C::ST_LD + k];
        }
      if (t < 128) {
        if (t < rows_here && p.ld) p.ld[row0 + t] += p.inv ? -s_ld[t] : s_ld[t];
        s_ld[t] = 0.f;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(x_empty(xs));
      if (!early_scatter && tile + gridDim.x < num_tiles) scatter_x2(tile + gridDim.x, tcount + 1);
    }
    if (p.y_meta) {
      run_max = warp_max(run_max);
      if (lane == 0) meta_amax(p.y_meta, run_max);
    }
    if (t == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
