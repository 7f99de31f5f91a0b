// K3f -- fused conditioner of an AffineCoupling (reference src/flows/realnvp.jl:57-110, src/flows/utils.jl:71-100).
//
// One persistent CTA per SM carries 128-sample tiles through BOTH conditioner networks (s and t), all three Dense
// layers each, and the coupling arithmetic, without a hidden activation ever visiting HBM as an operand:
//
//   X tile (fp32, one 1-D bulk copy) -> x2 = X[:, idx2] split into fp16 hi/lo planes (smem, UMMA K-major)
//   L1: acc[128 x 64] = x2 W1[chunk]^T        A from smem,  B (weights) streamed by TMA through a 3 x 16 KB ring
//       epilogue: bias + leakyrelu + hi/lo split -> TMEM (tcgen05.st) as the A operand of L2   (+ stash, see below)
//   L2: acc[128 x 64] = h1 W2[chunk, K chunks]^T  A from TMEM (tcgen05.mma .ts form); an accumulation chain covers two K
//       chunks of 64, chains are summed in registers with round-to-nearest adds (the tensor core truncates its fp32
//       accumulator; rz_compensation removes the remaining bias)
//       epilogue: bias + leakyrelu + split -> TMEM chunk buffer = A operand of L3
//   L3: acc[128 x 32] += h2[chunk] W3[:, chunk]^T, drained per slab; s = tanh(.), t = (.)
//   coupling: y1 = exp(s) x1 + t, logdet += sum(s)   (inverse direction: x1 = (y1 - t) exp(-s), logdet -= sum(s))
//
// TMEM (512 columns): [0,128) h1 hi | [128,256) h1 lo | [256,384) two 64-column accumulators (one per team) |
//                     [384,416) h2 chunk hi | [416,448) h2 chunk lo | [448,512) two 32-column L3 accumulators
// fp16 operands in TMEM are packed two per 32-bit column (element k of row m: lane m, column k/2, half k%2).
//
// Warp roles (640 threads):
//   warp 0      producer: TMA weight stages in schedule order, bulk copy of the next X tile
//   warp 1      MMA issuer: walks the schedule with warp-uniform control flow, one elected lane issues (no divisions in the loop)
//   warps 2-3   tile warps: build the NEXT tile's x2 planes as soon as this tile's first-Dense MMAs are complete, and do the
//               coupling arithmetic of the finished tile from the staged exp(+-s) / t tiles (rows re-read after an L2 prefetch)
//   warps 4-19  epilogue in TWO TEAMS of eight (TMEM lane quarter = warp % 4).  Team t owns accumulator t and the hidden chunks
//               j with j % 2 == t; a thread owns 32 of the chunk's 64 columns.  The schedule alternates the teams' chains, so
//               one team's activation / split pass runs while the tensor pipe works for the other team.
// The item schedule (fused_build_schedule, host side) is a STREAM across networks and tiles: the next network's first Dense is
// issued in this network's tail, the last third-Dense slabs of a network are drained inside the next one.  Hand-overs:
// mbarriers between the async proxies (TMA, tcgen05) and the warps; two named barriers per tile between the epilogue and the
// tile warps for the staged network outputs.  Measurements and the reasons for each piece: profiles/r2_fused_fwd_notes.md.
//
// What still goes to HBM is the stash the backward pass consumes (same formats as the layer-by-layer path, so
// tc_mlp_backward works unchanged): x2 planes, h1 / h2 planes + sign bits per network, s (fp32), and the new state.  Calls
// that no backward pass follows (sampling, logpdf, value-only objectives) skip every stash store (p.no_stash).
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
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr_a, uint32_t taddr_b, uint32_t (&va)[16], uint32_t (&vb)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(va[0]), "=r"(va[1]), "=r"(va[2]), "=r"(va[3]), "=r"(va[4]), "=r"(va[5]), "=r"(va[6]), "=r"(va[7]),
        "=r"(va[8]), "=r"(va[9]), "=r"(va[10]), "=r"(va[11]), "=r"(va[12]), "=r"(va[13]), "=r"(va[14]), "=r"(va[15])
      : "r"(taddr_a)
      : "memory");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(vb[0]), "=r"(vb[1]), "=r"(vb[2]), "=r"(vb[3]), "=r"(vb[4]), "=r"(vb[5]), "=r"(vb[6]), "=r"(vb[7]),
        "=r"(vb[8]), "=r"(vb[9]), "=r"(vb[10]), "=r"(vb[11]), "=r"(vb[12]), "=r"(vb[13]), "=r"(vb[14]), "=r"(vb[15])
      : "r"(taddr_b)
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
// producer side of a named barrier: non-blocking arrival (the consumers bar.sync on the same id / count)
__device__ __forceinline__ void epi_bar_arrive(int id, int nthreads) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
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
  int no_stash;             // 1: no backward pass will follow -- skip every stash store (hidden planes, sign bits, x2 planes, s)
  long long* dbg;           // optional clock64 timeline of CTA 0: [role][512]
  FusedNet net[2];
  // streaming variant: per-network item schedules (fused_build_schedule), MMA / producer order and epilogue order
  int n_seq_m, n_seq_e;
  int slab;                 // K chunks (of 64) per tensor-memory accumulation chain of the second Dense: 1 or 2
  int n_hoist;              // first-Dense chunks issued during the previous network's tail
  uint8_t seq_m[48], seq_e[48];
};

#define NF_FDBG(role, idx) do { if (p.dbg && blockIdx.x == 0 && (idx) < 512) p.dbg[(role) * 512 + (idx)] = clock64(); } while (0)

struct FusedFwdMaps {
  CUtensorMap w[2][3];      // weight planes, K-major, box {64, 64 | 32, 1}
  CUtensorMap x2;           // stash of the x2 planes: box {64, 128, 1}
  CUtensorMap h[2][2];      // stash of the hidden planes [net][layer]: box {16, 32, 1}, SWIZZLE_32B
  CUtensorMap w128[2][2];   // wide variant: first / second Dense weight planes with box {64, 128, 1}
  CUtensorMap h64[2][2];    // stash of the hidden planes, box {32, 32, 1}, SWIZZLE_64B (two-team kernel: one warp's piece)
};

struct FusedCfg {
  static constexpr int STAGES = 3;
  static constexpr int STAGE = 16384;            // weight slab: 64 rows x 128 B, hi plane then lo plane (+8192)
  static constexpr int XS = 32768;               // one X tile: 128 rows x d floats (d <= 64)
  static constexpr int X2_PLANE = 16384;         // 128 rows x 128 B
  static constexpr int ST_LD = 33;               // padded row stride of the s / t staging tiles (floats)
  static constexpr int OFF_W = 0;
  static constexpr int OFF_X = OFF_W + STAGES * STAGE;
  static constexpr int OFF_X2 = OFF_X + XS;      // one X buffer: released once the conditioner input is built (the coupling pass re-reads global)
  static constexpr int OFF_S = OFF_X2 + 2 * X2_PLANE;
  static constexpr int OFF_T = OFF_S + 128 * ST_LD * 4;
  // staging of the hidden-activation stash: 16 warps x {hi, lo} x [32 rows x 64 B] (SWIZZLE_64B), left by TMA stores
  static constexpr int OFF_STG = (OFF_T + 128 * ST_LD * 4 + 1023) / 1024 * 1024;
  static constexpr int OFF_BIAS = OFF_STG + 16 * 4096;             // [2 nets][256 + 256 + 32] floats
  static constexpr int OFF_LD = OFF_BIAS + 2 * 544 * 4;           // [4 column groups][128 rows] partial logdets (summed in a fixed order)
  static constexpr int OFF_BAR = OFF_LD + 2048;
  static constexpr int N_BARS = 2 * STAGES + 2 /*x full/empty*/ + 2 /*x2 ready/free*/ + 4 /*tfull/tempty*/ + 4 /*tfull3/tempty3*/ +
                                4 /*h1 ready*/ + 2 /*h2 ready/free*/;
  static constexpr int OFF_POS = OFF_BAR + 8 * N_BARS + 16;   // pos[64], pos2[64]
  static constexpr int SMEM = OFF_POS + 512;
  // warp 0 producer, 1 MMA issuer, 2-3 tile warps (conditioner input, coupling arithmetic), 4..19 epilogue
  static constexpr int EPI0 = 128;
  static constexpr int THREADS = EPI0 + 512;
  // TMEM columns
  static constexpr int TM_H1HI = 0, TM_H1LO = 128, TM_ACC = 256, TM_H2HI = 384, TM_H2LO = 416, TM_ACC3 = 448;
};
static_assert(FusedCfg::SMEM <= 232448, "fused coupling: shared memory budget");

// Per-network schedule of the streaming kernel.  Items (one byte each):
//   (j << 2) | k   second-Dense K chunk k of hidden chunk j of THIS network; | 0x10: first of its tensor-memory accumulation
//                  chain (waits for the accumulator), | 0x20: last of the chain (commits it; the epilogue drains it)
//   0x60 | j       first-Dense chunk j of the NEXT network (hoisted: its epilogue then overlaps this network's tail);
//   0x40 | j       first-Dense chunk j of this network (chunks the first chains do not read)
//   0x80 | j       third-Dense slab j of this network;  0xA0 | j: of the PREVIOUS network (it wrapped past the end)
// Chunk j belongs to team j & 1 (accumulator j & 1).  The two teams' K loops alternate with team 0 two slabs ahead, the next
// network's first Dense follows the last slab (its epilogue overwrites the h1 planes, so every reader must have been issued),
// and the third-Dense slab of a chunk comes `delay` items after the chunk's last slab (its operand needs the chunk's
// activation pass first).  The epilogue's copy drains the third Dense later still (the two 32-column accumulators decouple
// them).  Third-Dense slabs stay in chunk order in both copies, everything else is in the same order in both.
inline int fused_build_schedule(uint8_t* seq, int cap, int nch, int delay, int slab, int n_hoist, int lead = 1) {
  std::vector<uint8_t> base;
  int jt[2] = {0, 1}, kt[2] = {0, 0};
  // one accumulation chain (`slab` consecutive K chunks of one hidden chunk) per turn
  auto take = [&](int tm) -> bool {
    if (jt[tm] >= nch) return false;
    for (int u = 0; u < slab && kt[tm] < nch; ++u, ++kt[tm]) {
      const bool first = u == 0, last = u == slab - 1 || kt[tm] == nch - 1;
      base.push_back((uint8_t)((jt[tm] << 2) | kt[tm] | (first ? 0x10 : 0) | (last ? 0x20 : 0)));
    }
    if (kt[tm] == nch) { kt[tm] = 0; jt[tm] += 2; }
    return true;
  };
  // The first chains of both teams read only the hoisted h1 chunks; the first Dense of the remaining chunks (0x40 | j, THIS
  // network) follows them, so its epilogues overlap those chains instead of delaying them.
  int turns = 1;
  take(0);
  if (slab == 1 || lead > 1) { take(0); ++turns; }
  int tm = 1;
  bool placed = n_hoist >= nch;
  while (jt[0] < nch || jt[1] < nch) {
    if (!take(tm)) take(tm ^ 1);
    tm ^= 1;
    if (!placed && ++turns >= (slab == 1 ? 3 : 2)) {
      for (int j = n_hoist; j < nch; ++j) base.push_back((uint8_t)(0x40 | j));
      placed = true;
    }
  }
  for (int j = 0; j < std::min(n_hoist, nch); ++j) base.push_back((uint8_t)(0x60 | j));
  const int len = (int)base.size();
  // after[i]: third-Dense items that follow base[i]
  std::vector<std::vector<uint8_t>> after(len);
  for (int pass = 0; pass < 2; ++pass)            // wrapped (previous network) items first: they precede this network's
    for (int j = 0; j < nch; ++j) {
      int last = 0;
      for (int i = 0; i < len; ++i) if (!(base[i] & 0xC0) && (base[i] & 0xF) == ((j << 2) | (nch - 1))) last = i;
      const int pos = last + delay;
      if (pass == 0 && pos >= len) after[std::min(pos - len, len - 1)].push_back((uint8_t)(0xA0 | j));
      if (pass == 1 && pos < len) after[pos].push_back((uint8_t)(0x80 | j));
    }
  int n = 0;
  for (int i = 0; i < len; ++i) {
    if (n < cap) seq[n] = base[i];
    ++n;
    for (uint8_t e : after[i]) { if (n < cap) seq[n] = e; ++n; }
  }
  return n;
}

__device__ __forceinline__ uint32_t fused_act_split(const uint32_t (&v)[16], float ds, const float* __restrict__ bias, uint32_t (&hi)[8],
                                                    uint32_t (&lo)[8]) {
  uint32_t bits = 0;
#pragma unroll
  for (int q = 0; q < 16; q += 2) {
    float a = fmaf(__uint_as_float(v[q]), ds, bias[q]);
    float b = fmaf(__uint_as_float(v[q + 1]), ds, bias[q + 1]);
    bits |= (a > 0.f ? 1u : 0u) << q;
    bits |= (b > 0.f ? 1u : 0u) << (q + 1);
    a = fmaxf(a, 0.01f * a); b = fmaxf(b, 0.01f * b);
    split_pair(a, b, hi[q >> 1], lo[q >> 1]);
  }
  return bits;
}


// One warp leaves its [32 rows x 16 columns] piece of a hidden-activation chunk: registers -> the warp's staging tile
// (SWIZZLE_32B pattern, conflict-free 16-byte stores) -> one TMA store per plane.  Rows past the batch are clipped by the map.
__device__ __forceinline__ void fused_stash_store(uint8_t* stg, const CUtensorMap* map, const uint32_t (&hi)[8], const uint32_t (&lo)[8],
                                                  int lane, int col0, int row_base, bool two_planes, bool wait = true) {
  // one ELECTED lane waits / issues (elect.sync keeps the bulk-tensor instructions on the uniform datapath; a `lane == 0`
  // branch costs an election loop per instruction)
  if (wait) {
    if (elect_one_sync()) tma_store_wait_read();   // the previous stores of this warp have read the tile
    __syncwarp();
  }
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

// One warp of the two-team kernel leaves its [32 rows x 32 columns] piece of a hidden-activation chunk: registers -> the warp's
// staging tile (SWIZZLE_64B pattern: 16-byte chunk index ^ ((row >> 1) & 3), conflict-free 16-byte stores) -> one TMA store per
// plane.  Rows past the batch are clipped by the map.
__device__ __forceinline__ void fused_stash_store64(uint8_t* stg, const CUtensorMap* map, const uint32_t (&hi0)[8], const uint32_t (&lo0)[8],
                                                    const uint32_t (&hi1)[8], const uint32_t (&lo1)[8], int lane, int col0, int row_base,
                                                    bool two_planes, int dbg = 0) {
  if (!(dbg & (1 << 18))) {
  if (elect_one_sync()) tma_store_wait_read();     // the warp's previous stores have read the tile
  __syncwarp();
  }
  const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
  uint8_t* rowp = stg + lane * 64;
  *reinterpret_cast<uint4*>(rowp + ((0u ^ sw) << 4)) = make_uint4(hi0[0], hi0[1], hi0[2], hi0[3]);
  *reinterpret_cast<uint4*>(rowp + ((1u ^ sw) << 4)) = make_uint4(hi0[4], hi0[5], hi0[6], hi0[7]);
  *reinterpret_cast<uint4*>(rowp + ((2u ^ sw) << 4)) = make_uint4(hi1[0], hi1[1], hi1[2], hi1[3]);
  *reinterpret_cast<uint4*>(rowp + ((3u ^ sw) << 4)) = make_uint4(hi1[4], hi1[5], hi1[6], hi1[7]);
  if (two_planes) {
    *reinterpret_cast<uint4*>(rowp + 2048 + ((0u ^ sw) << 4)) = make_uint4(lo0[0], lo0[1], lo0[2], lo0[3]);
    *reinterpret_cast<uint4*>(rowp + 2048 + ((1u ^ sw) << 4)) = make_uint4(lo0[4], lo0[5], lo0[6], lo0[7]);
    *reinterpret_cast<uint4*>(rowp + 2048 + ((2u ^ sw) << 4)) = make_uint4(lo1[0], lo1[1], lo1[2], lo1[3]);
    *reinterpret_cast<uint4*>(rowp + 2048 + ((3u ^ sw) << 4)) = make_uint4(lo1[4], lo1[5], lo1[6], lo1[7]);
  }
  if (!(dbg & (1 << 17))) fence_proxy_async();
  __syncwarp();
  if (!(dbg & (1 << 16)) && elect_one_sync()) {
    tma_store_3d(map, smem_u32(stg), col0, row_base, 0);
    if (two_planes) tma_store_3d(map, smem_u32(stg) + 2048, col0, row_base, 1);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(FusedCfg::THREADS, 1)
fused_affine_fwd_kernel(const __grid_constant__ FusedFwdMaps maps, const __grid_constant__ FusedFwdParams p) {
  using C = FusedCfg;
  constexpr int S = C::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  const uint32_t bars = base + C::OFF_BAR;
  auto w_full = [&](int s) { return bars + 8u * s; };
  auto w_empty = [&](int s) { return bars + 8u * (S + s); };
  const uint32_t x_full = bars + 8u * (2 * S), x_empty = bars + 8u * (2 * S + 1);
  const uint32_t x2_ready = bars + 8u * (2 * S + 2), x2_free = bars + 8u * (2 * S + 3);
  auto tfull = [&](int a) { return bars + 8u * (2 * S + 4 + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * S + 6 + a); };
  auto tfull3 = [&](int a) { return bars + 8u * (2 * S + 8 + a); };
  auto tempty3 = [&](int a) { return bars + 8u * (2 * S + 10 + a); };
  auto h1_ready = [&](int k) { return bars + 8u * (2 * S + 12 + k); };
  const uint32_t h2_ready = bars + 8u * (2 * S + 16), h2_free = bars + 8u * (2 * S + 17);
  const uint32_t tmem_slot = bars + 8u * C::N_BARS;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + C::OFF_BAR + 8 * C::N_BARS);
  float* s_bias = reinterpret_cast<float*>(smem_raw + C::OFF_BIAS);
  float* s_ld = reinterpret_cast<float*>(smem_raw + C::OFF_LD);
  float* s_S = reinterpret_cast<float*>(smem_raw + C::OFF_S);
  float* s_T = reinterpret_cast<float*>(smem_raw + C::OFF_T);
  int* s_pos = reinterpret_cast<int*>(smem_raw + C::OFF_POS);
  int* s_pos2 = s_pos + 64;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int64_t num_tiles = (p.n + 127) / 128;
  const int nch = p.nch;
  const int n_seq_m = p.n_seq_m, n_seq_e = p.n_seq_e;
  // this CTA's tiles: blockIdx.x, + gridDim.x, ...; network instance q = 2 * (tile count) + network
  const int my_tiles = (int)((num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const int nq = 2 * my_tiles;

  // ---- scales (every thread derives the same values from the same device scalars) ----
  const float amax_x = __uint_as_float(reinterpret_cast<const unsigned int*>(p.x_meta)[1]);
  const float s_x2 = pow2_scale(amax_x);

  if (warp == 0 && lane == 0) {
    for (int nt = 0; nt < 2; ++nt)
      for (int l = 0; l < 3; ++l) tma_prefetch_desc(&maps.w[nt][l]);
    tma_prefetch_desc(&maps.x2);
    for (int s = 0; s < S; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    mbar_init(x_full, 1); mbar_init(x_empty, 1);
    mbar_init(x2_ready, 1); mbar_init(x2_free, 1);
    // accumulator a, the hidden chunks j with j % 2 == a and their operand planes belong to team a (eight warps)
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 8); mbar_init(tfull3(a), 1); mbar_init(tempty3(a), 16); }
    for (int k = 0; k < 4; ++k) mbar_init(h1_ready(k), 8);
    mbar_init(h2_ready, 8); mbar_init(h2_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (threadIdx.x >= C::EPI0) {
    const int t = threadIdx.x - C::EPI0;
    // zero the x2 planes once (padding columns stay zero), the logdet staging, and fetch the biases
    for (int i = t; i < 2 * C::X2_PLANE / 16; i += 512) reinterpret_cast<uint4*>(smem_raw + C::OFF_X2)[i] = make_uint4(0, 0, 0, 0);
    if (t < p.d) { s_pos[t] = p.pos[t]; s_pos2[t] = p.pos2[t]; }
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
    // warp-uniform: every lane walks the schedule, one elected lane arms the barrier and issues the copies
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
        if (p.terms > 1) tma_load_3d(st + 8192, map, w_full(s), c0, c1, 1);
      }
      __syncwarp();
      ++it;
    };
    load_x(blockIdx.x);
    for (int j = 0; j < p.n_hoist; ++j) load_w(&maps.w[0][0], 0, j * 64, 8192);
    for (int q = 0; q <= nq; ++q) {
      const int nt = q & 1;
      // the X buffer is released as soon as a tile's conditioner input has been built: the next tile's rows land early
      if (q < nq && nt == 0 && (q >> 1) + 1 < my_tiles) load_x(blockIdx.x + (int64_t)((q >> 1) + 1) * gridDim.x);
      for (int i = 0; i < n_seq_m; ++i) {
        const int e = p.seq_m[i];
        if (e & 0x80) {
          const bool prev = (e & 0x20) != 0;
          if (prev ? q == 0 : q == nq) continue;
          load_w(&maps.w[prev ? nt ^ 1 : nt][2], (e & 3) * 64, 0, 4096);
        } else if (q == nq) {
          continue;                                 // drain pass: only the wrapped third-Dense slabs of the last network
        } else if (e & 0x40) {
          const bool next = (e & 0x20) != 0;
          if (next && q == nq - 1) continue;
          load_w(&maps.w[next ? nt ^ 1 : nt][0], 0, (e & 3) * 64, 8192);
        } else {
          load_w(&maps.w[nt][1], (e & 3) * 64, ((e >> 2) & 3) * 64, 8192);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // The whole warp walks the schedule (uniform control flow, barrier waits by every lane) and ONE elected lane issues:
    // with `elect.sync` the compiler keeps descriptors in uniform registers and emits back-to-back UTCHMMA; a plain
    // `lane == 0` branch costs an election loop (~60 cycles) per MMA, more than a 128 x 64 x 16 MMA takes to execute.
    constexpr uint32_t idesc64 = make_idesc(128, 64, 0, 0);
    constexpr uint32_t idesc32 = make_idesc(128, 32, 0, 0);
    uint32_t it = 0, sl_a = 0, sl_b = 0, sl3 = 0, cc = 0;
    int s = 0;                 // weight ring stage of item `it` and its phase (kept incrementally: no division in this loop)
    uint32_t wph = 0;
    auto next_stage = [&]() { ++it; if (++s == S) { s = 0; wph ^= 1; } };
    const bool t3 = p.terms > 1;
    const uint32_t x2a = base + C::OFF_X2;
    const int kk1 = p.kk1;
    // first Dense, chunk j of network nt_l of this CTA's tile number tile_l: A = x2 planes in shared memory
    auto issue_l1 = [&](int nt_l, int tile_l, int j) {
      const uint32_t st = base + C::OFF_W + s * C::STAGE;
      const uint64_t b_hi = make_smem_desc(st, 16, 1024), b_lo = make_smem_desc(st + 8192, 16, 1024);
      const uint32_t acc = j & 1;
      if (nt_l == 0 && j == 0) mbar_wait(x2_ready, tile_l & 1);
      mbar_wait3(tempty(acc), ((acc ? sl_b : sl_a) & 1) ^ 1, w_full(s), wph, 0, 0);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + C::TM_ACC + acc * 64;
      const uint64_t a_hi = make_smem_desc(x2a, 16, 1024), a_lo = make_smem_desc(x2a + C::X2_PLANE, 16, 1024);
      if (elect_one_sync()) {
        uint32_t accum = 0;
        if (t3) {
          for (int kk = 0; kk < kk1; ++kk) { umma_f16(d_tmem, a_lo + ((kk * 32) >> 4), b_hi + ((kk * 32) >> 4), idesc64, accum); accum = 1; }
          for (int kk = 0; kk < kk1; ++kk) umma_f16(d_tmem, a_hi + ((kk * 32) >> 4), b_lo + ((kk * 32) >> 4), idesc64, 1u);
        }
        for (int kk = 0; kk < kk1; ++kk) { umma_f16(d_tmem, a_hi + ((kk * 32) >> 4), b_hi + ((kk * 32) >> 4), idesc64, accum); accum = 1; }
        umma_commit(w_empty(s));
        umma_commit(tfull(acc));
        if (nt_l == 1 && j == nch - 1) umma_commit(x2_free);   // this tile's planes may be overwritten
      }
      __syncwarp();
      if (acc) ++sl_b; else ++sl_a;
      next_stage();
    };
    for (int j = 0; j < p.n_hoist; ++j) issue_l1(0, 0, j);
    for (int q = 0; q <= nq; ++q) {
      for (int i = 0; i < n_seq_m; ++i) {
        const int e = p.seq_m[i];
        if (e & 0x80) {
          // ---- third Dense, slab j: A = the chunk's h2 planes in tensor memory ----
          if ((e & 0x20) ? q == 0 : q == nq) continue;
              const uint32_t st = base + C::OFF_W + s * C::STAGE;
          const uint64_t b_hi = make_smem_desc(st, 16, 1024), b_lo = make_smem_desc(st + 8192, 16, 1024);
          const uint32_t acc = sl3 & 1;
          NF_FDBG(0, 4 * it);
          mbar_wait3(tempty3(acc), ((sl3 >> 1) & 1) ^ 1, w_full(s), wph, h2_ready, cc & 1);
          NF_FDBG(0, 4 * it + 1);
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
          NF_FDBG(0, 4 * it + 2);
          ++sl3; ++cc; next_stage();
        } else if (q == nq) {
          continue;
        } else if (e & 0x40) {
          const int ql = (e & 0x20) ? q + 1 : q;
          if (ql == nq) continue;
          NF_FDBG(0, 4 * it);
          issue_l1(ql & 1, ql >> 1, e & 3);
          NF_FDBG(0, 4 * (it - 1) + 2);
        } else {
          // ---- second Dense, slab k of chunk j: A = h1 planes in tensor memory ----
              const uint32_t st = base + C::OFF_W + s * C::STAGE;
          const uint64_t b_hi = make_smem_desc(st, 16, 1024), b_lo = make_smem_desc(st + 8192, 16, 1024);
          const int j = (e >> 2) & 3, k = e & 3;
          const uint32_t acc = j & 1;
          // an accumulation chain covers p.slab K chunks: the first waits for the accumulator, the last commits it
          const bool chain_first = (e & 0x10) != 0, chain_last = (e & 0x20) != 0;
          NF_FDBG(0, 4 * it);
          // independent try_waits in one loop overlap their latencies (a completed try_wait still costs ~200 cycles)
          mbar_wait3(chain_first ? tempty(acc) : w_full(s), chain_first ? ((acc ? sl_b : sl_a) & 1) ^ 1 : wph, w_full(s), wph,
                     j == 0 ? h1_ready(k) : 0u, q & 1);
          NF_FDBG(0, 4 * it + 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + C::TM_ACC + acc * 64;
          const uint32_t a_hi = tmem_base + C::TM_H1HI + k * 32, a_lo = tmem_base + C::TM_H1LO + k * 32;
          if (elect_one_sync()) {
            const uint32_t cont = chain_first ? 0u : 1u;
            if (t3) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16_ts(d_tmem, a_lo + kk * 8, b_hi + ((kk * 32) >> 4), idesc64, kk > 0 ? 1u : cont);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16_ts(d_tmem, a_hi + kk * 8, b_lo + ((kk * 32) >> 4), idesc64, 1u);
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_f16_ts(d_tmem, a_hi + kk * 8, b_hi + ((kk * 32) >> 4), idesc64, (t3 || kk > 0) ? 1u : cont);
            umma_commit(w_empty(s));
            if (chain_last) umma_commit(tfull(acc));
          }
          __syncwarp();
          NF_FDBG(0, 4 * it + 2);
          if (chain_last) { if (acc) ++sl_b; else ++sl_a; }
          next_stage();
        }
      }
    }
  } else if (warp < 4) {
    // =========================== tile warps (2-3): conditioner input of the next tile, coupling arithmetic of this one ===========================
    // x2 = X[:, idx2] * s_x2 -> hi / lo planes (UMMA K-major, SWIZZLE_128B) of the NEXT tile while the epilogue warps still
    // work on this one; then, once both networks' outputs of this tile are staged, y1 = exp(s) x1 + t and the logdet.
    // The epilogue warps never leave the MMA stream.
    const int tx = threadIdx.x - 64;               // 0 .. 63
    const int d = p.d, dq = p.d >> 2;              // float4 per row
    const float4* Xs4 = reinterpret_cast<const float4*>(smem_raw + C::OFF_X);
    // element idx = tx + 64 i  ->  (row, float4 column).  When 64 % dq == 0 a thread keeps ONE float4 column for the whole
    // tile, so its four mask lookups (and everything derived from them) leave the loops.
    const bool fixed_col = (64 % dq) == 0;
    const int qc0 = tx % dq, r00 = tx / dq, rstep = 64 / dq;
    int k1f[4], k2f[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { k1f[u] = s_pos[4 * qc0 + u]; k2f[u] = s_pos2[4 * qc0 + u]; }
    float run_max = 0.f;
    auto build_x2 = [&](int tc) {
      const int64_t r0 = (blockIdx.x + (int64_t)tc * gridDim.x) * 128;
      const int rows_s = (int)((p.n - r0) < 128 ? (p.n - r0) : 128);
      // (one lane polls: thirty-two lanes polling one barrier serialise their ~200-cycle try_waits)
      if (lane == 0) {
        mbar_wait(x_full, tc & 1);
        if (tc > 0) mbar_wait(x2_free, (tc - 1) & 1);   // the previous tile's first-Dense MMAs are complete
        if (tx == 0) tma_store_wait_read();              // ... and its x2 stash store has read the planes
      }
      epi_bar_sync(3, 64);
      auto put4 = [&](const float4 x, int r, int k0, int k1, int k2, int k3) {
        const float xv[4] = {x.x, x.y, x.z, x.w};
        const int kk[4] = {k0, k1, k2, k3};
        const uint32_t rbase = (uint32_t)r * 128u, rsw = (uint32_t)r & 7u;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = kk[u];
          if (k >= 0) {
            const float v = xv[u] * s_x2;
            const __half h = __float2half_rn(v);
            const __half l = __float2half_rn(v - __half2float(h));
            const uint32_t off = rbase + ((((uint32_t)k >> 3) ^ rsw) << 4) + ((uint32_t)k & 7u) * 2u;
            *reinterpret_cast<__half*>(smem_raw + C::OFF_X2 + off) = h;
            *reinterpret_cast<__half*>(smem_raw + C::OFF_X2 + C::X2_PLANE + off) = l;
          }
        }
      };
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (fixed_col) {
        // (compact loops on purpose: the kernel's code is far larger than the instruction cache, straight-line code starves)
#pragma unroll 2
        for (int r = r00; r < 128; r += rstep) put4(r < rows_s ? Xs4[r * dq + qc0] : zero4, r, k2f[0], k2f[1], k2f[2], k2f[3]);
      } else {
#pragma unroll 1
        for (int idx = tx; idx < 128 * dq; idx += 64) {
          const int r = idx / dq, qc = idx - r * dq;
          put4(r < rows_s ? Xs4[idx] : zero4, r, s_pos2[4 * qc], s_pos2[4 * qc + 1], s_pos2[4 * qc + 2], s_pos2[4 * qc + 3]);
        }
      }
      fence_proxy_async();
      epi_bar_sync(3, 64);
      if (tx == 0) {
        mbar_arrive(x2_ready);
        mbar_arrive(x_empty);                      // the X buffer is free again: the next tile's rows may land
        if (!p.no_stash) {
          tma_store_3d(&maps.x2, base + C::OFF_X2, 0, (int)r0, 0);     // stash of the x2 planes for the weight-gradient kernel
          if (p.terms > 1) tma_store_3d(&maps.x2, base + C::OFF_X2 + C::X2_PLANE, 0, (int)r0, 1);
        }
      }
    };
    build_x2(0);
    for (int tc = 0; tc < my_tiles; ++tc) {
      if (tx == 0) NF_FDBG(3, 8 * tc);
      if (tc + 1 < my_tiles) build_x2(tc + 1);
      if (tx == 0) NF_FDBG(3, 8 * tc + 1);
      const int64_t row0 = (blockIdx.x + (int64_t)tc * gridDim.x) * 128;
      const int rows_here = (int)((p.n - row0) < 128 ? (p.n - row0) : 128);
      const float4* xrow0 = reinterpret_cast<const float4*>(p.Xin + row0 * d);
      float4* yrow0 = reinterpret_cast<float4*>(p.Xout + row0 * d);
      // the tile's rows were read a whole tile ago: bring them back into L2 while the second network finishes
      for (int o = tx * 128; o < rows_here * d * 4; o += 64 * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(xrow0) + o));
      epi_bar_sync(1, 576);                          // both networks' outputs of this tile are staged (512 epilogue threads arrive)
      if (tx == 0) NF_FDBG(3, 8 * tc + 2);
      // ---- coupling arithmetic: coalesced float4 passes ----
      auto couple4 = [&](const float4 x, int r, int k0, int k1, int k2, int k3) -> float4 {
        float y[4] = {x.x, x.y, x.z, x.w};
        const int kk[4] = {k0, k1, k2, k3};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int k = kk[h];
          if (k >= 0) {
            const float ev = s_S[r * C::ST_LD + k], tv = s_T[r * C::ST_LD + k];    // exp(+-s), t
            y[h] = p.inv ? (y[h] - tv) * ev : fmaf(ev, y[h], tv);
          }
          run_max = fmaxf(run_max, fabsf(y[h]));
        }
        return make_float4(y[0], y[1], y[2], y[3]);
      };
      if (fixed_col) {
        // four rows per round: the loads of a round are in flight together (L2 hits after the prefetch above)
#pragma unroll 1
        for (int r = r00; r < rows_here; r += 4 * rstep) {
          float4 xv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { const int ru = r + u * rstep; if (ru < rows_here) xv[u] = __ldcg(xrow0 + ru * dq + qc0); }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int ru = r + u * rstep;
            if (ru < rows_here) yrow0[ru * dq + qc0] = couple4(xv[u], ru, k1f[0], k1f[1], k1f[2], k1f[3]);
          }
        }
      } else {
#pragma unroll 1
        for (int idx = tx; idx < rows_here * dq; idx += 64) {
          const int r = idx / dq, qc = idx - r * dq;
          yrow0[idx] = couple4(__ldcg(xrow0 + idx), r, s_pos[4 * qc], s_pos[4 * qc + 1], s_pos[4 * qc + 2], s_pos[4 * qc + 3]);
        }
      }
      if (tx == 0) NF_FDBG(3, 8 * tc + 7);
      for (int r = tx; r < rows_here; r += 64) {
        const float sum_s = (s_ld[r] + s_ld[128 + r]) + (s_ld[256 + r] + s_ld[384 + r]);
        if (p.ld) p.ld[row0 + r] += p.inv ? -sum_s : sum_s;
      }
      epi_bar_arrive(2, 576);                        // the staging tiles may take the next tile's outputs
      if (tx == 0) NF_FDBG(3, 8 * tc + 3);
    }
    if (p.y_meta) {
      run_max = warp_max(run_max);
      if (lane == 0) meta_amax(p.y_meta, run_max);
    }
    if (tx == 0) tma_store_wait_all();
  } else {
    // =========================== epilogue warps ===========================
    const int t = threadIdx.x - C::EPI0;           // 0 .. 511
    const int ew = warp - 4;                       // 0 .. 15
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may touch
    const int team = ew >> 3;                      // owns accumulator `team` and the hidden chunks j with j % 2 == team
    const int cg = (ew >> 2) & 1;                  // 32-column half of the team's 64-column chunk
    const int g = ew >> 2;                         // 8-column group of the 32 third-Dense outputs
    const int rloc = quarter * 32 + lane;          // row of the tile this thread owns in TMEM
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int h_ld = p.h_ld;
    const int bits_ld = h_ld >> 4;                 // 16-bit words per row of the sign bits
    uint8_t* stg_warp = smem_raw + C::OFF_STG + ew * 4096;
    const bool two = p.terms > 1;
    uint32_t nb = 0, sl3 = 0;
    // per-network constants
    float ds1[2], ds2[2], ds3[2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const FusedNet& N = p.net[nt];
      const float bound1 = amax_x * N.w_sc[0][1] + N.w_sc[0][3];
      const float s_h1 = pow2_scale(bound1 * 1.001f);
      const float bound2 = bound1 * N.w_sc[1][1] + N.w_sc[1][3];
      const float s_h2 = pow2_scale(bound2 * 1.001f);
      const float d1 = 1.f / (s_x2 * N.w_sc[0][0]), d2 = 1.f / (s_h1 * N.w_sc[1][0]), d3 = 1.f / (s_h2 * N.w_sc[2][0]);
      ds1[nt] = fmaf(d1, p.rz[0], d1) * s_h1; ds2[nt] = fmaf(d2, p.rz[1], d2) * s_h2; ds3[nt] = fmaf(d3, p.rz[2], d3);
      if (blockIdx.x == 0 && t == 0) {
        N.h_meta[0][0] = s_h1; N.h_meta[0][1] = bound1;
        N.h_meta[1][0] = s_h2; N.h_meta[1][1] = bound2;
        if (nt == 0) { p.x2_meta[0] = s_x2; p.x2_meta[1] = amax_x; }
      }
    }
    const uint32_t acc_addr = tmem_base + lane_off + C::TM_ACC + team * 64 + cg * 32;
    // wait for the team's accumulator, load this thread's 32 columns, release the accumulator
    auto drain = [&](uint32_t (&v0)[16], uint32_t (&v1)[16]) {
      if (t == team * 256) NF_FDBG(1 + team, 4 * nb);
      if (lane == 0) mbar_wait_relaxed(tfull(team), nb & 1);
      __syncwarp();
      if (t == team * 256) NF_FDBG(1 + team, 4 * nb + 1);
      tc_fence_after();
      tmem_ld16x2(acc_addr, acc_addr + 16, v0, v1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(team));
      if (t == team * 256) NF_FDBG(1 + team, 4 * nb + 2);
      ++nb;
    };
    // stash of one chunk piece: planes through the pair's staging tile, sign bits straight to global memory
    auto stash = [&](int nt, int layer, int j, int64_t row0, const uint32_t (&hi0)[8], const uint32_t (&lo0)[8], const uint32_t (&hi1)[8],
                     const uint32_t (&lo1)[8], uint32_t bits) {
      if (p.dbg_flags & 1) return;
      fused_stash_store64(stg_warp, &maps.h64[nt][layer], hi0, lo0, hi1, lo1, lane, j * 64 + cg * 32, (int)row0 + quarter * 32, two, p.dbg_flags);
      if (row0 + rloc < p.n && !(p.dbg_flags & (1 << 19)))
        *reinterpret_cast<uint32_t*>(p.net[nt].h_bits[layer] + (row0 + rloc) * bits_ld + ((j * 64 + cg * 32) >> 4)) = bits;
    };
    // first-Dense epilogue of chunk j of network nt: this thread's 32 columns of h1 -> TMEM operand planes + stash
    auto l1_epilogue = [&](int nt, int j, int64_t row0) {
      // (two 16-column halves one after the other: this may run while a second-Dense chain's sums are live in registers)
      if (t == team * 256) NF_FDBG(1 + team, 4 * nb);
      if (lane == 0) mbar_wait_relaxed(tfull(team), nb & 1);
      __syncwarp();
      if (t == team * 256) NF_FDBG(1 + team, 4 * nb + 1);
      tc_fence_after();
      const int col = j * 64 + cg * 32;
      const float* b1 = s_bias + nt * 544;
      uint32_t hi0[8], lo0[8], hi1[8], lo1[8], bits;
      {
        uint32_t v[16];
        tmem_ld16(acc_addr, v);
        bits = fused_act_split(v, ds1[nt], b1 + col, hi0, lo0);
      }
      {
        uint32_t v[16];
        tmem_ld16(acc_addr + 16, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(team));
        bits |= fused_act_split(v, ds1[nt], b1 + col + 16, hi1, lo1) << 16;
      }
      if (t == team * 256) NF_FDBG(1 + team, 4 * nb + 2);
      ++nb;
      const uint32_t h1a = tmem_base + lane_off + j * 32 + cg * 16;
      tmem_st8(h1a + C::TM_H1HI, hi0); tmem_st8(h1a + C::TM_H1HI + 8, hi1);
      tmem_st8(h1a + C::TM_H1LO, lo0); tmem_st8(h1a + C::TM_H1LO + 8, lo1);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(h1_ready(j));
      if (t == team * 256) NF_FDBG(1 + team, 4 * (nb - 1) + 3);
      stash(nt, 0, j, row0, hi0, lo0, hi1, lo1, bits);
    };
    float racc[32], racc3[8];
#pragma unroll
    for (int q8 = 0; q8 < 8; ++q8) racc3[q8] = 0.f;
    // outputs of network instance qq are complete in racc3: s (after tanh) / t -> staging tile for the tile warps
    auto net_end = [&](int qq) {
      const int nt = qq & 1, tc = qq >> 1;
      const float* b3 = s_bias + nt * 544 + 512;
      const int64_t row = (blockIdx.x + (int64_t)tc * gridDim.x) * 128 + rloc;
      float a[8];
#pragma unroll
      for (int q8 = 0; q8 < 8; ++q8) { a[q8] = fmaf(racc3[q8], ds3[nt], b3[g * 8 + q8]); racc3[q8] = 0.f; }
      if (nt == 0) {
        float part = 0.f;
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) { a[q8] = tanhf(a[q8]); if (g * 8 + q8 < p.c) part += a[q8]; }
        if (tc > 0) {                                // the tile warps are done with the previous tile's staging
          epi_bar_sync(2, 576);
        }
        // the tile warps get exp(+-s) (their pass is then one FMA per element); s itself goes straight to the stash
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) s_S[rloc * C::ST_LD + g * 8 + q8] = expf(p.inv ? -a[q8] : a[q8]);
        s_ld[g * 128 + rloc] = part;
      } else {
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) s_T[rloc * C::ST_LD + g * 8 + q8] = a[q8];
        epi_bar_arrive(1, 576);
      }
      float* o = p.net[nt].out;
      if (o && row < p.n) {
        const int c = p.c;
        o += row * c + g * 8;
        if ((c & 3) == 0 && g * 8 + 8 <= c) {
          reinterpret_cast<float4*>(o)[0] = make_float4(a[0], a[1], a[2], a[3]);
          reinterpret_cast<float4*>(o)[1] = make_float4(a[4], a[5], a[6], a[7]);
        } else {
#pragma unroll
          for (int q8 = 0; q8 < 8; ++q8) if (g * 8 + q8 < c) o[q8] = a[q8];
        }
      }
    };
    for (int j = team; j < p.n_hoist; j += 2) l1_epilogue(0, j, (int64_t)blockIdx.x * 128);
    for (int q = 0; q <= nq; ++q) {
      const int nt = q & 1;
      const int64_t row0 = (blockIdx.x + (int64_t)(q >> 1) * gridDim.x) * 128;
      for (int i = 0; i < n_seq_e; ++i) {
        const int e = p.seq_e[i];
        if (e & 0x80) {
          // ---- third Dense slab (both teams): this thread's 8 of the 32 outputs ----
          const bool prev = (e & 0x20) != 0;
          if (prev ? q == 0 : q == nq) continue;
          const uint32_t acc = sl3 & 1;
          if (lane == 0) mbar_wait_relaxed(tfull3(acc), (sl3 >> 1) & 1);
          __syncwarp();
          tc_fence_after();
          uint32_t v[8];
          tmem_ld8(tmem_base + lane_off + C::TM_ACC3 + acc * 32 + g * 8, v);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty3(acc));
          ++sl3;
#pragma unroll
          for (int q8 = 0; q8 < 8; ++q8) racc3[q8] += __uint_as_float(v[q8]);
          if ((e & 3) == nch - 1) net_end(prev ? q - 1 : q);
          continue;
        }
        if (q == nq) continue;
        if (e & 0x40) {
          const int ql = (e & 0x20) ? q + 1 : q;
          if (ql == nq || ((e & 1) != team)) continue;
          l1_epilogue(ql & 1, e & 3, (blockIdx.x + (int64_t)(ql >> 1) * gridDim.x) * 128);
          continue;
        }
        const int j = (e >> 2) & 3, k = e & 3;
        if ((j & 1) != team) continue;
        if (!(e & 0x20)) continue;                                     // the accumulation chain continues with the next K chunk
        // ---- second Dense: chains summed in registers (round to nearest) ----
        uint32_t v0[16], v1[16];
        drain(v0, v1);
        if (k < p.slab) {
#pragma unroll
          for (int c = 0; c < 16; ++c) { racc[c] = __uint_as_float(v0[c]); racc[16 + c] = __uint_as_float(v1[c]); }
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) { racc[c] += __uint_as_float(v0[c]); racc[16 + c] += __uint_as_float(v1[c]); }
        }
        if (k == nch - 1) {
          // chunk j of h2: activation + split, then it becomes the A operand of the third Dense
          const int col = j * 64 + cg * 32;
          const float* b2 = s_bias + nt * 544 + 256;
          uint32_t hi0[8], lo0[8], hi1[8], lo1[8];
#pragma unroll
          for (int c = 0; c < 16; ++c) { v0[c] = __float_as_uint(racc[c]); v1[c] = __float_as_uint(racc[16 + c]); }
          const uint32_t bits0 = fused_act_split(v0, ds2[nt], b2 + col, hi0, lo0);
          const uint32_t bits1 = fused_act_split(v1, ds2[nt], b2 + col + 16, hi1, lo1);
          // chunk order: the third-Dense MMAs of the previous chunk (the other team's) have read the buffer.  The tensor pipe
          // completes in order and this chunk's last slab was issued after the third Dense of the chunk two back, so the
          // barrier is at most one phase behind the one waited for here.
          const uint32_t cc = (uint32_t)q * (uint32_t)nch + (uint32_t)j;
          if (cc > 0 && lane == 0) mbar_wait(h2_free, (cc - 1) & 1);
          __syncwarp();
          tc_fence_after();
          const uint32_t h2a = tmem_base + lane_off + cg * 16;
          tmem_st8(h2a + C::TM_H2HI, hi0); tmem_st8(h2a + C::TM_H2HI + 8, hi1);
          tmem_st8(h2a + C::TM_H2LO, lo0); tmem_st8(h2a + C::TM_H2LO + 8, lo1);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(h2_ready);
          if (t == team * 256) NF_FDBG(1 + team, 4 * (nb - 1) + 3);
          stash(nt, 1, j, row0, hi0, lo0, hi1, lo1, bits0 | (bits1 << 16));
        }
      }
    }
    // every thread that issued bulk stores (stash pieces) waits for them before the CTA's shared memory goes away
    __syncwarp();
    if (elect_one_sync()) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
