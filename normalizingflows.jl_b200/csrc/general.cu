// Layered execution of flows with coupling layers: orchestrates per-layer kernels over sample chunks.
//
//   ELBO  (reference src/objectives/elbo.jl:65-70,89-92):  x0 -> [layers, last first] -> y ; head ; backward sweep
//   LOGLIK(reference src/objectives/loglikelihood.jl:26-33): y -> [inverse layers, first first] -> x0 ; head ; backward
//
// Conditioner MLPs (fnn, reference src/flows/utils.jl:71-100) run either on the CUDA-core GEMM
// (NF_MMA_SIMT: Float64 flows and validation) or on the tcgen05 GEMM of tc_gemm.cu (Float32).
#include "general.hpp"
#include "kernels_coupling.cuh"
#include "tc_gemm.hpp"
#include "base_dense.cuh"

namespace nf {

namespace {

struct LayerBufs {
  void* act0 = nullptr;                       // gathered conditioner input x2 [n, cbar]
  std::vector<std::vector<void*>> acts;       // acts[m][i-1] = output of Dense i-1 (post activation), i = 1..n_dense
};

struct Chunk {
  int64_t n = 0;
  bool stash = false;
  std::vector<void*> X;                       // stash: L+1 states; else input + 2 ping-pong buffers
  void* xin(int state) const { return stash ? X[state] : (state == 0 ? X[0] : X[1 + ((state - 1) & 1)]); }
  void* xout(int state) const { return stash ? X[state + 1] : X[1 + (state & 1)]; }
  void* ld = nullptr;
  void* G = nullptr;
  void* gld = nullptr;                        // optional per-sample d/dlogdet (two-phase API)
  std::vector<LayerBufs> lb;                  // stash: one per layer; else a single shared set at [0]
  void* ga[4] = {nullptr, nullptr, nullptr, nullptr};  // backward temporaries (width = max MLP width)
  std::vector<float*> xmeta;                  // tcgen05 mode: {scale, max|X|} slot per state (bound for the fp16 operand scale)
  int n_states = 0;
};

struct GenState {
  Chunk stash_chunk;     // kept alive between nf_forward_stash and nf_backward
  bool has_stash = false;
};

inline bool is_coupling(int kind) { return kind == NF_AFFINE_COUPLING || kind == NF_SPLINE_COUPLING; }
// planar / radial layers composed with couplings (reference src/flows/utils.jl:23-26 puts no restriction on Ls) run as "segments":
// a maximal run of consecutive planar / radial / shift layers goes through the fused elementwise kernel in one launch
inline bool is_seg_kind(int kind) { return kind == NF_PLANAR || kind == NF_RADIAL || kind == NF_SHIFT; }
inline bool is_seg_core(int kind) { return kind == NF_PLANAR || kind == NF_RADIAL; }
// [lo, hi] = the maximal run of segment-kind layers containing layer li, if that run holds a planar / radial layer (else lo = hi = -1)
inline void segment_of(const Flow& f, int li, int* lo, int* hi) {
  *lo = *hi = -1;
  if (!is_seg_kind(f.layers[li].kind)) return;
  int a = li, b = li;
  while (a > 0 && is_seg_kind(f.layers[a - 1].kind)) --a;
  while (b + 1 < (int)f.layers.size() && is_seg_kind(f.layers[b + 1].kind)) ++b;
  bool core = false;
  for (int i = a; i <= b; ++i) core |= is_seg_core(f.layers[i].kind);
  if (core) { *lo = a; *hi = b; }
}

int max_width(const Flow& f) {
  int w = 1;
  for (auto& L : f.layers)
    for (auto& m : L.mlps)
      for (int v : m.dims) w = std::max(w, v);
  return w;
}

// bytes of one activation buffer holding `width` features for n samples
size_t act_bytes(const Flow& f, int64_t n, int width, bool final_out) {
  if (f.mma_mode == NF_MMA_SIMT || final_out) return (size_t)n * width * f.esize();
  return tc_act_bytes(n, width);
}

size_t per_sample_bytes(const Flow& f, bool stash) {
  const size_t es = f.esize();
  const int L = (int)f.layers.size();
  size_t b = 0;
  b += (size_t)(stash ? L + 1 : 3) * f.dim * es;   // X
  b += 3 * es;                                      // ld, gld, lq0 (full-covariance base)
  b += (size_t)f.dim * es;                          // G
  size_t layer_max = 0, layer_sum = 0;
  for (auto& Ld : f.layers) {
    if (!is_coupling(Ld.kind)) continue;
    size_t lbts = act_bytes(f, 1024, (int)Ld.idx2.size(), false);
    for (auto& m : Ld.mlps)
      for (int i = 1; i < (int)m.dims.size(); ++i) lbts += act_bytes(f, 1024, m.dims[i], i + 1 == (int)m.dims.size());
    layer_sum += lbts;
    layer_max = std::max(layer_max, lbts);
  }
  b += (stash ? layer_sum : layer_max) / 1024 + 1;
  b += 4 * act_bytes(f, 1024, max_width(f), false) / 1024 + 4 * (size_t)max_width(f) * es;
  return b;
}

int alloc_chunk(Flow& f, Chunk& c, int64_t n, bool stash, const void* x0_alias) {
  const size_t es = f.esize();
  const int L = (int)f.layers.size();
  c.n = n; c.stash = stash;
  c.n_states = stash ? L + 1 : 3;
  c.X.assign(c.n_states, nullptr);
  for (int i = 0; i < c.n_states; ++i) {
    if (i == 0 && x0_alias) { c.X[0] = const_cast<void*>(x0_alias); continue; }
    c.X[i] = f.ws_alloc((size_t)n * f.dim * es);
    if (!c.X[i]) return NF_ERR_OOM;
  }
  c.ld = f.ws_alloc((size_t)n * es);
  c.G = f.ws_alloc((size_t)n * f.dim * es);
  if (!c.ld || !c.G) return NF_ERR_OOM;
  const int nlb = stash ? L : 1;
  c.lb.assign(nlb, LayerBufs());
  if (stash) {
    for (int li = 0; li < L; ++li) {
      const LayerDesc& Ld = f.layers[li];
      if (!is_coupling(Ld.kind)) continue;
      LayerBufs& b = c.lb[li];
      b.act0 = f.ws_alloc(act_bytes(f, n, (int)Ld.idx2.size(), false));
      if (!b.act0) return NF_ERR_OOM;
      b.acts.resize(Ld.mlps.size());
      for (size_t m = 0; m < Ld.mlps.size(); ++m) {
        const MLPDesc& md = Ld.mlps[m];
        for (int i = 1; i < (int)md.dims.size(); ++i) {
          void* p = f.ws_alloc(act_bytes(f, n, md.dims[i], i + 1 == (int)md.dims.size()));
          if (!p) return NF_ERR_OOM;
          b.acts[m].push_back(p);
        }
      }
    }
  } else {
    // one shared set sized for the widest layer
    LayerBufs& b = c.lb[0];
    int cb = 1; size_t nm = 0; std::vector<std::vector<int>> w;
    for (auto& Ld : f.layers) {
      if (!is_coupling(Ld.kind)) continue;
      cb = std::max(cb, (int)Ld.idx2.size());
      nm = std::max(nm, Ld.mlps.size());
    }
    b.act0 = f.ws_alloc(act_bytes(f, n, cb, false));
    if (!b.act0) return NF_ERR_OOM;
    b.acts.resize(nm);
    for (size_t m = 0; m < nm; ++m) {
      size_t nd = 0;
      for (auto& Ld : f.layers) if (m < Ld.mlps.size()) nd = std::max(nd, Ld.mlps[m].dims.size() - 1);
      for (size_t i = 1; i <= nd; ++i) {
        size_t bytes = 0;
        for (auto& Ld : f.layers)
          if (m < Ld.mlps.size() && i < Ld.mlps[m].dims.size()) {
            bytes = std::max(bytes, act_bytes(f, n, Ld.mlps[m].dims[i], false));
            bytes = std::max(bytes, act_bytes(f, n, Ld.mlps[m].dims[i], true));
          }
        void* p = f.ws_alloc(bytes);
        if (!p) return NF_ERR_OOM;
        b.acts[m].push_back(p);
      }
    }
  }
  return NF_OK;
}

int alloc_backward_tmps(Flow& f, Chunk& c) {
  const int mw = max_width(f);
  for (int i = 0; i < 4; ++i) {
    size_t bytes = std::max(act_bytes(f, c.n, mw, false), (size_t)c.n * mw * f.esize());
    c.ga[i] = f.ws_alloc(bytes);
    if (!c.ga[i]) return NF_ERR_OOM;
  }
  return NF_OK;
}

// ---------------------------------------------------------------------------------------------
// CUDA-core MLP forward / backward
// ---------------------------------------------------------------------------------------------
template <typename T, bool TA, bool TB, typename Epi>
int launch_simt_gemm(Flow& f, const T* A, int64_t lda, const T* B, int64_t ldb, int64_t M, int N, int64_t K, int64_t ksplit,
                     const Epi& epi) {
  if (ksplit <= 0) ksplit = K;
  dim3 grid((unsigned)ceil_div(M, 64), (unsigned)ceil_div(N, 64), (unsigned)ceil_div(K, ksplit));
  simt_gemm_kernel<T, TA, TB, Epi><<<grid, 256, 0, f.stream>>>(A, lda, B, ldb, M, N, K, ksplit, epi);
  NF_LAUNCH_CHECK();
  return NF_OK;
}

template <typename T>
int simt_mlp_forward(Flow& f, const MLPDesc& md, const T* theta, int64_t n, const T* act0, std::vector<void*>& acts) {
  const T* in = act0;
  const int nd = md.n_dense();
  for (int i = 0; i < nd; ++i) {
    T* out = (T*)acts[i];
    const int act = (i + 1 < nd) ? ACT_LRELU : (md.out_act ? ACT_TANH : ACT_NONE);
    EpiBiasAct<T> epi{out, md.dims[i + 1], theta + md.b_off[i], act};
    NF_TRY((launch_simt_gemm<T, false, false>(f, in, md.dims[i], theta + md.w_off[i], md.dims[i + 1], n, md.dims[i + 1],
                                              md.dims[i], 0, epi)));
    in = out;
  }
  return NF_OK;
}

// g_last: gradient w.r.t. the pre-activation of the last Dense [n, out].  Accumulates parameter
// gradient sums into gsum (theta order) and scatter-adds the conditioner-input gradient into G[:, idx2].
template <typename T>
int simt_mlp_backward(Flow& f, const MLPDesc& md, const T* theta, int64_t n, const T* act0, std::vector<void*>& acts,
                      T* g_last, T* tmp0, T* tmp1, T* G, int d, const int* d_idx2, double* gsum, bool last_bias_done = false) {
  const int nd = md.n_dense();
  T* g = g_last;
  for (int i = nd - 1; i >= 0; --i) {
    const T* in = (i == 0) ? act0 : (const T*)acts[i - 1];
    const int kin = md.dims[i], kout = md.dims[i + 1];
    if (!(last_bias_done && i == nd - 1)) {
      const int64_t rpb = 4096;
      colsum_atomic_kernel<T><<<(unsigned)ceil_div(n, rpb), 256, 0, f.stream>>>(g, n, kout, rpb, gsum + md.b_off[i]);
      NF_LAUNCH_CHECK();
    }
    {  // wgrad: dWt[k][o] = sum_n in[n][k] g[n][o]
      EpiAtomicDouble<T> epi{gsum + md.w_off[i], kout};
      NF_TRY((launch_simt_gemm<T, true, false>(f, in, kin, g, kout, kin, kout, n, 2048, epi)));
    }
    if (i > 0) {  // dgrad into hidden: gin[n][k] = (sum_o g[n][o] Wt[k][o]) * lrelu'(h[n][k])
      T* gin = (g == tmp0) ? tmp1 : tmp0;
      EpiMaskLrelu<T> epi{gin, kin, in, kin};
      NF_TRY((launch_simt_gemm<T, false, true>(f, g, kout, theta + md.w_off[i], kout, n, kin, kout, 0, epi)));
      g = gin;
    } else if (G) {
      EpiScatterAdd<T> epi{G, d, d_idx2};
      NF_TRY((launch_simt_gemm<T, false, true>(f, g, kout, theta + md.w_off[i], kout, n, kin, kout, 0, epi)));
    }
  }
  return NF_OK;
}

// ---------------------------------------------------------------------------------------------
// one coupling layer, forward (INV = false: y = T(x)) or inverse direction
// ---------------------------------------------------------------------------------------------
// Launch shape of the spline kernels: threads per block (= pairs per tile), ring depth and grid such that
// stages * tile fits in shared memory with several blocks resident per SM.
struct RqsLaunch { int threads, stages; size_t smem; unsigned grid; };
template <typename T>
RqsLaunch rqs_launch_shape(int K, int64_t pairs) {
  const size_t row = (size_t)(3 * K - 1) * sizeof(T);
  int thr = 128;
  while (thr > 32 && 2 * thr * row > (size_t)96 * 1024) thr >>= 1;
  const size_t tile = thr * row;
  // one tile of prefetch is enough (a tile's arithmetic takes far longer than the HBM latency); the shared memory is
  // better spent on resident warps to hide the dependent cumsum / division chains
  int stages = 2;
  if (const char* e = getenv("NFCUDA_RQS_STAGES")) stages = std::max(2, std::min(rq::kMaxStages, atoi(e)));
  const size_t smem = stages * tile;
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)216 * 1024) / (smem + 1024)));
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(pairs, thr), (int64_t)per_sm * kNumSMs));
  return RqsLaunch{thr, stages, smem, grid};
}

template <typename T, bool INV>
int coupling_apply(Flow& f, const LayerDesc& Ld, LayerBufs& b, const T* theta, int64_t n, const T* Xin, T* Xout, T* ld,
                   int32_t* bins, const float* amax_in, float* amax_out, bool stash = true) {
  const int d = f.dim, c = (int)Ld.idx1.size(), cbar = (int)Ld.idx2.size();
  const bool tc = f.mma_mode != NF_MMA_SIMT;
  if (Ld.kind == NF_SHIFT || Ld.kind == NF_SCALE) {
    diag_apply_kernel<T, INV><<<(unsigned)std::min<int64_t>(ceil_div(n * d, 256), 16 * kNumSMs), 256, 0, f.stream>>>(
        Xin, theta + Ld.theta_off, Ld.kind == NF_SCALE ? 1 : 0, d, n, Xout, ld, amax_out);
    NF_LAUNCH_CHECK();
    return NF_OK;
  }
  if (tc && amax_in && tc_fused_affine_ok(f, Ld)) {
    // both conditioners, all Dense layers and the coupling arithmetic in one launch (fused_coupling.cuh)
    return tc_affine_forward_fused(f, Ld, n, (const float*)Xin, (float*)Xout, (float*)ld, b.act0, b.acts, amax_in, amax_out, INV, stash);
  }
  if (tc) {
    NF_TRY(tc_gather_split(f, (const float*)Xin, d, Ld.d_idx2, cbar, n, b.act0, amax_in));
    for (size_t m = 0; m < Ld.mlps.size(); ++m) NF_TRY(tc_mlp_forward(f, Ld, (int)m, n, b.act0, b.acts[m]));
  } else {
    gather_cols_kernel<T><<<(unsigned)ceil_div(n * cbar, 256), 256, 0, f.stream>>>(Xin, d, Ld.d_idx2, cbar, n, (T*)b.act0);
    NF_LAUNCH_CHECK();
    for (size_t m = 0; m < Ld.mlps.size(); ++m) NF_TRY(simt_mlp_forward<T>(f, Ld.mlps[m], theta, n, (const T*)b.act0, b.acts[m]));
  }
  if (Ld.kind == NF_AFFINE_COUPLING) {
    const T* Sv = (const T*)b.acts[0].back();
    const T* Tv = (const T*)b.acts[1].back();
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 8), 8 * kNumSMs);
    f.prof.begin("affine_apply", f.stream);
    if (d <= 32) affine_apply_rows_kernel<T, INV, 1><<<grid, 256, 0, f.stream>>>(Xin, Sv, Tv, Ld.d_pos, c, d, n, Xout, ld, amax_out);
    else if (d <= 64) affine_apply_rows_kernel<T, INV, 2><<<grid, 256, 0, f.stream>>>(Xin, Sv, Tv, Ld.d_pos, c, d, n, Xout, ld, amax_out);
    else if (d <= 128) affine_apply_rows_kernel<T, INV, 4><<<grid, 256, 0, f.stream>>>(Xin, Sv, Tv, Ld.d_pos, c, d, n, Xout, ld, amax_out);
    else if (d <= 256) affine_apply_rows_kernel<T, INV, 8><<<grid, 256, 0, f.stream>>>(Xin, Sv, Tv, Ld.d_pos, c, d, n, Xout, ld, amax_out);
    else
      affine_apply_kernel<T, INV><<<(unsigned)std::min<int64_t>(ceil_div(n * d, 256), 16 * kNumSMs), 256, 0, f.stream>>>(
          Xin, Sv, Tv, Ld.d_pos, c, d, n, Xout, ld, amax_out);
    f.prof.end(f.stream);
  } else {
    const RqsLaunch rl = rqs_launch_shape<T>(Ld.K, n * c);
    const T* raw = (const T*)b.acts[0].back();
#define NF_RQS_APPLY(KM)                                                                                              \
    do {                                                                                                                \
      auto kern = rqs_apply_kernel<T, KM, INV>;                                                                         \
      if (rl.smem > 48 * 1024) NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rl.smem)); \
      kern<<<rl.grid, rl.threads, rl.smem, f.stream>>>(Xin, raw, Ld.d_idx1, Ld.d_pos, c, d, Ld.K, (T)Ld.B, n, Xout, ld, bins, amax_out, rl.stages); \
    } while (0)
    f.prof.begin("rqs_apply", f.stream);
    if (Ld.K <= 8) NF_RQS_APPLY(8);
    else if (Ld.K <= 10) NF_RQS_APPLY(10);
    else if (Ld.K <= 16) NF_RQS_APPLY(16);
    else NF_RQS_APPLY(64);
    f.prof.end(f.stream);
#undef NF_RQS_APPLY
  }
  NF_LAUNCH_CHECK();
  return NF_OK;
}

// need_input_grad = false for the layer whose input is the batch itself (base draws / data): the gradient w.r.t. its
// conditioner input is never used, so the first-Dense dgrad (GEMM + scatter) is skipped.
template <typename T, bool INV>
int coupling_backward(Flow& f, const LayerDesc& Ld, LayerBufs& b, const T* theta, Chunk& c, const T* Xin, const T* Xout,
                      bool need_input_grad) {
  const int d = f.dim, cc = (int)Ld.idx1.size();
  const int64_t n = c.n;
  const bool tc = f.mma_mode != NF_MMA_SIMT;
  T* G = (T*)c.G;
  const T* gld = (const T*)c.gld;
  T* gA = (T*)c.ga[0];
  T* gB = (T*)c.ga[1];
  T* gC = (T*)c.ga[2];
  if (Ld.kind == NF_SHIFT || Ld.kind == NF_SCALE) {
    const int64_t rpb = 4096;
    diag_bwd_kernel<T, INV><<<(unsigned)ceil_div(n, rpb), 256, 0, f.stream>>>(G, INV ? Xout : Xin, theta + Ld.theta_off, gld,
                                                                            Ld.kind == NF_SCALE ? 1 : 0, d, n, rpb, f.d_gsum + Ld.theta_off);
    NF_LAUNCH_CHECK();
    return NF_OK;
  }
  if (Ld.kind == NF_AFFINE_COUPLING) {
    // gA <- d/d(pre-tanh s), gB <- d/dt
    float* mS = tc ? tc_alloc_meta(f) : nullptr;
    float* mT = tc ? tc_alloc_meta(f) : nullptr;
    if (tc) NF_REQUIRE(mS && mT, "tcgen05 path: out of tensor metadata slots");
    affine_bwd_kernel<T, INV><<<(unsigned)std::min<int64_t>(ceil_div(n * cc, 256), 16 * kNumSMs), 256, 0, f.stream>>>(
        G, INV ? Xout : Xin, (const T*)b.acts[0].back(), gld, Ld.d_idx1, cc, d, n, gA, gB, mS, mT);
    NF_LAUNCH_CHECK();
    if (tc) {
      NF_TRY(tc_mlp_backward(f, Ld, 0, n, b.act0, b.acts[0], (float*)gA, mS, c.ga[2], c.ga[3], need_input_grad ? (float*)G : nullptr, f.d_gsum));
      NF_TRY(tc_mlp_backward(f, Ld, 1, n, b.act0, b.acts[1], (float*)gB, mT, c.ga[2], c.ga[3], need_input_grad ? (float*)G : nullptr, f.d_gsum));
    } else {
      // s network: temporaries gC + (gA after it has been consumed is NOT safe) -> use a dedicated pair
      NF_TRY(simt_mlp_backward<T>(f, Ld.mlps[0], theta, n, (const T*)b.act0, b.acts[0], gA, gC, gA, need_input_grad ? G : nullptr, d, Ld.d_idx2, f.d_gsum));
      NF_TRY(simt_mlp_backward<T>(f, Ld.mlps[1], theta, n, (const T*)b.act0, b.acts[1], gB, gC, gB, need_input_grad ? G : nullptr, d, Ld.d_idx2, f.d_gsum));
    }
  } else {
    float* mR = tc ? tc_alloc_meta(f) : nullptr;
    if (tc) NF_REQUIRE(mR, "tcgen05 path: out of tensor metadata slots");
    bool fused_bias = false;
    const RqsLaunch rl = rqs_launch_shape<T>(Ld.K, n * cc);
    // bias gradient of the conditioner's last Dense = column sums of graw: folded into the kernel when a tile is whole samples
    const MLPDesc& md = Ld.mlps[0];
    fused_bias = (rl.threads % cc == 0) && (cc * (3 * Ld.K - 1) <= 4 * rl.threads);
    double* cs = fused_bias ? f.d_gsum + md.b_off[md.n_dense() - 1] : nullptr;
    const T* raw = (const T*)b.acts[0].back();
    // tcgen05 path: the gradient w.r.t. the conditioner output leaves the spline kernel as the split planes the GEMMs read
    // (RqsPlanesOut: estimate pass over every 64th tile, speculative pass, commit-or-redo) instead of an fp32 matrix plus a
    // split pass.  nf_set_option("rqs_planes", 0) / NFCUDA_RQS_PLANES=0: the fp32 form (A/B switch); 2 forces the redo pass (tests).
    const int ncol = cc * (3 * Ld.K - 1);
    const bool planes = tc && g_opt_rqs_planes != 0 && fused_bias && (ncol % 2 == 0) && sizeof(T) == 4;
    RqsPlanesOut<T> po{};
    if (planes) {
      TcPlanesOut P;
      NF_TRY(tc_planes_out(f, c.ga[2], n, ncol, &P));
      float* est = tc_alloc_meta(f);
      NF_REQUIRE(est, "tcgen05 path: out of tensor metadata slots");
      po.hi = (__half*)P.hi; po.plane_elems = P.plane_elems; po.ld = P.ld; po.meta = P.meta; po.est = est;
      po.gside = gA; po.tile_step = 64;
      po.headroom = g_opt_rqs_planes == 2 ? 1.0e-6f : kRqsSpecHeadroom;
      NF_REQUIRE(P.ld <= 4 * rl.threads, "spline backward: planes wider than four column pairs per thread pair");
    }
#define NF_RQS_BWD(KM)                                                                                                \
    do {                                                                                                                \
      auto kern = rqs_bwd_kernel<T, KM, INV, 0>;                                                                        \
      auto kern1 = rqs_bwd_kernel<T, KM, INV, 1>;                                                                       \
      auto kern2 = rqs_bwd_kernel<T, KM, INV, 2>;                                                                       \
      auto kern3 = rqs_bwd_kernel<T, KM, INV, 3>;                                                                       \
      if (rl.smem > 48 * 1024) {                                                                                        \
        NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rl.smem));                 \
        NF_CUDA(cudaFuncSetAttribute(kern1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rl.smem));                \
        NF_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rl.smem));                \
        NF_CUDA(cudaFuncSetAttribute(kern3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rl.smem));                \
      }                                                                                                                 \
      if (!planes) {                                                                                                    \
        kern<<<rl.grid, rl.threads, rl.smem, f.stream>>>(G, Xin, raw, gld, Ld.d_idx1, cc, d, Ld.K, (T)Ld.B, n, gA, mR, cs, rl.stages, po); \
      } else {                                                                                                          \
        const int64_t tiles_all = (n * cc) / rl.threads;                                                                \
        const unsigned g1 = (unsigned)std::max<int64_t>(1, std::min<int64_t>(rl.grid, ceil_div(tiles_all, po.tile_step))); \
        po.mode = 1;                                                                                                    \
        kern1<<<g1, rl.threads, rl.smem, f.stream>>>(G, Xin, raw, gld, Ld.d_idx1, cc, d, Ld.K, (T)Ld.B, n, gA, nullptr, nullptr, rl.stages, po); \
        po.mode = 2;                                                                                                    \
        kern2<<<rl.grid, rl.threads, rl.smem, f.stream>>>(G, Xin, raw, gld, Ld.d_idx1, cc, d, Ld.K, (T)Ld.B, n, gA, po.meta, cs, rl.stages, po); \
        nf::g_launch_count += 2;      /* three launches, one NF_LAUNCH_CHECK below */                                   \
        po.mode = 3;                                                                                                    \
        kern3<<<rl.grid, rl.threads, rl.smem, f.stream>>>(G, Xin, raw, gld, Ld.d_idx1, cc, d, Ld.K, (T)Ld.B, n, gA, po.meta, cs, rl.stages, po); \
      }                                                                                                                 \
    } while (0)
    f.prof.begin("rqs_bwd", f.stream);
    if (Ld.K <= 8) NF_RQS_BWD(8);
    else if (Ld.K <= 10) NF_RQS_BWD(10);
    else if (Ld.K <= 16) NF_RQS_BWD(16);
    else NF_RQS_BWD(64);
    f.prof.end(f.stream);
#undef NF_RQS_BWD
    NF_LAUNCH_CHECK();
    if (tc) NF_TRY(tc_mlp_backward(f, Ld, 0, n, b.act0, b.acts[0], planes ? nullptr : (float*)gA, mR, c.ga[2], c.ga[3], need_input_grad ? (float*)G : nullptr, f.d_gsum, fused_bias));
    else NF_TRY(simt_mlp_backward<T>(f, Ld.mlps[0], theta, n, (const T*)b.act0, b.acts[0], gA, gC, gB, need_input_grad ? G : nullptr, d, Ld.d_idx2, f.d_gsum, fused_bias));
  }
  return NF_OK;
}

template <typename T>
int check_supported(const Flow& f, int op) {
  bool seg = false;
  for (auto& L : f.layers) {
    if (is_seg_core(L.kind)) { seg = true; continue; }
    if (!is_coupling(L.kind) && L.kind != NF_SHIFT && L.kind != NF_SCALE) {
      set_error("flows mixing Hamiltonian (LeapFrog / momentum) layers with coupling layers are not supported in this build");
      return NF_ERR_UNSUPPORTED;
    }
  }
  if (seg && f.dim > 64) {
    set_error("planar / radial layers inside a coupling flow support dim <= 64, got %d", f.dim);
    return NF_ERR_UNSUPPORTED;
  }
  return NF_OK;
}

// forward sweep of one chunk (application order: last layer first).  Leaves the result in c.X[last].
template <typename T>
int sweep_forward(Flow& f, Chunk& c, const T* theta, int32_t* bins, int64_t bins_chunk_off, int64_t N_total) {
  const int L = (int)f.layers.size();
  NF_CUDA(cudaMemsetAsync(c.ld, 0, (size_t)c.n * sizeof(T), f.stream));
  int state = 0;
  int64_t bins_layer_off = 0;   // bins layout [spline layer in application order][N_total][c]
  for (int li = L - 1; li >= 0; --li) {
    const LayerDesc& Ld = f.layers[li];
    int slo, shi;
    segment_of(f, li, &slo, &shi);
    if (slo >= 0) {
      // layers slo..shi (shi applied first) in one fused launch; the states in between are never materialised
      const int k = shi - slo + 1;
      const T* Xin = (const T*)c.xin(state);
      T* Xout = (T*)c.xout(state + k - 1);
      NF_TRY(ew_segment<T>(f, slo, k, theta, c.n, Xin, Xout, c.ld, false, nullptr, nullptr, nullptr));
      if (!c.xmeta.empty()) NF_TRY(tc_absmax(f, (const float*)Xout, c.n * f.dim, c.xmeta[state + k]));
      state += k;
      li = slo;
      continue;
    }
    LayerBufs& b = c.stash ? c.lb[li] : c.lb[0];
    const T* Xin = (const T*)c.xin(state);
    T* Xout = (T*)c.xout(state);
    int32_t* bl = nullptr;
    if (bins && Ld.kind == NF_SPLINE_COUPLING) {
      bl = bins + bins_layer_off + bins_chunk_off * (int64_t)Ld.idx1.size();
      bins_layer_off += N_total * (int64_t)Ld.idx1.size();
    }
    const bool track = !c.xmeta.empty();
    NF_TRY((coupling_apply<T, false>(f, Ld, b, theta, c.n, Xin, Xout, (T*)c.ld, bl, c.xmeta.empty() ? nullptr : c.xmeta[state],
                                     track ? c.xmeta[state + 1] : nullptr, c.stash)));
    if (!c.xmeta.empty() && !track) c.xmeta[state + 1] = nullptr;   // unknown bound: the next split measures it
    ++state;
  }
  return NF_OK;
}

// inverse sweep (first layer first).
template <typename T>
int sweep_inverse(Flow& f, Chunk& c, const T* theta) {
  const int L = (int)f.layers.size();
  NF_CUDA(cudaMemsetAsync(c.ld, 0, (size_t)c.n * sizeof(T), f.stream));
  int state = 0;
  for (int li = 0; li < L; ++li) {
    const LayerDesc& Ld = f.layers[li];
    int slo, shi;
    segment_of(f, li, &slo, &shi);
    if (slo >= 0) {      // li == slo: the inverse of the run applies slo first
      const int k = shi - slo + 1;
      const T* Xin = (const T*)c.xin(state);
      T* Xout = (T*)c.xout(state + k - 1);
      NF_TRY(ew_segment<T>(f, slo, k, theta, c.n, Xin, Xout, c.ld, false, nullptr, nullptr, nullptr, true));
      if (!c.xmeta.empty()) NF_TRY(tc_absmax(f, (const float*)Xout, c.n * f.dim, c.xmeta[state + k]));
      state += k;
      li = shi;
      continue;
    }
    LayerBufs& b = c.stash ? c.lb[li] : c.lb[0];
    const T* Xin = (const T*)c.xin(state);
    T* Xout = (T*)c.xout(state);
    const bool track = !c.xmeta.empty();
    NF_TRY((coupling_apply<T, true>(f, Ld, b, theta, c.n, Xin, Xout, (T*)c.ld, nullptr, c.xmeta.empty() ? nullptr : c.xmeta[state],
                                    track ? c.xmeta[state + 1] : nullptr, c.stash)));
    if (!c.xmeta.empty() && !track) c.xmeta[state + 1] = nullptr;
    ++state;
  }
  return NF_OK;
}

template <typename T>
int sweep_backward_fwd(Flow& f, Chunk& c, const T* theta) {   // backward of the forward sweep
  const int L = (int)f.layers.size();
  int state = L;
  for (int li = 0; li < L; ++li) {   // layer 0 was applied last
    int slo, shi;
    segment_of(f, li, &slo, &shi);
    if (slo >= 0) {
      const int k = shi - slo + 1;     // li == slo here: the run is entered at its lowest index
      NF_TRY(ew_segment<T>(f, slo, k, theta, c.n, c.X[state - k], nullptr, nullptr, true, c.G, c.gld, f.d_gsum));
      state -= k;
      li = shi;
      continue;
    }
    NF_TRY((coupling_backward<T, false>(f, f.layers[li], c.lb[li], theta, c, (const T*)c.X[state - 1], (const T*)c.X[state], state - 1 > 0)));
    --state;
  }
  return NF_OK;
}

template <typename T>
int sweep_backward_inv(Flow& f, Chunk& c, const T* theta) {   // backward of the inverse sweep
  const int L = (int)f.layers.size();
  int state = L;
  for (int li = L - 1; li >= 0; --li) {   // layer L-1's inverse was applied last
    int slo, shi;
    segment_of(f, li, &slo, &shi);
    if (slo >= 0) {      // li == shi
      const int k = shi - slo + 1;
      NF_TRY(ew_segment<T>(f, slo, k, theta, c.n, c.X[state - k], nullptr, nullptr, true, c.G, c.gld, f.d_gsum, true));
      state -= k;
      li = slo;
      continue;
    }
    NF_TRY((coupling_backward<T, true>(f, f.layers[li], c.lb[li], theta, c, (const T*)c.X[state - 1], (const T*)c.X[state], state - 1 > 0)));
    --state;
  }
  return NF_OK;
}

template <typename T>
int run_typed(Flow& f, const GeneralJob& job) {
  NF_TRY(check_supported<T>(f, job.op));
  const T* theta = (const T*)job.theta_dev;
  const int d = f.dim, L = (int)f.layers.size();
  const int64_t N = job.N;
  const bool grad = job.want_grad;
  const bool stash = grad || job.op == OP_FORWARD_STASH;
  const int64_t Nc = std::min<int64_t>(N, f.chunk_N > 0 ? f.chunk_N : N);
  if (job.op == OP_FORWARD_STASH) NF_REQUIRE(Nc >= N, "nf_forward_stash: batch of %lld does not fit the workspace limit", (long long)N);
  if (job.op == OP_ELBO || job.op == OP_LOGLIK) NF_CUDA(cudaMemsetAsync(f.d_gsum, 0, (f.P + 1) * sizeof(double), f.stream));
  if (f.mma_mode != NF_MMA_SIMT) NF_TRY(tc_prepare_weights(f, (const float*)theta));
  const T* base = (f.base_is_standard || f.base_dense) ? nullptr : (const T*)f.d_base;
  const size_t ws_mark = f.ws.off;
  for (int64_t c0 = 0; c0 < N; c0 += Nc) {
    const int64_t n = std::min(Nc, N - c0);
    f.ws.off = ws_mark;
    if (c0 > 0 && f.in_ev_armed) {          // second half of a host-supplied batch: its copy ran beside the first half's compute
      NF_CUDA(cudaStreamWaitEvent(f.stream, f.in_ev, 0));
      f.in_ev_armed = false;
    }
    if (f.mma_mode != NF_MMA_SIMT && c0 > 0) NF_TRY(tc_begin_chunk(f));
    Chunk c;
    const T* in = job.in_dev ? (const T*)job.in_dev + c0 * d : nullptr;
    NF_TRY(alloc_chunk(f, c, n, stash, in));
    T* lq0 = nullptr;           // full-covariance base: per-sample log q0(x0) (and, for the log-likelihood, its gradient) from base_dense_kernel
    if (f.base_dense && (job.op == OP_ELBO || job.op == OP_LOGLIK)) {
      lq0 = (T*)f.ws_alloc((size_t)n * sizeof(T));
      if (!lq0) return NF_ERR_OOM;
    }
    if (!in) {
      base_sample_kernel<T><<<base_sample_grid(n, d, c0 + f.draw_row_offset), 256, 0, f.stream>>>((T*)c.X[0], base, d, n, job.seed, c0 + f.draw_row_offset, job.seed_iter_dev);
      NF_LAUNCH_CHECK();
      if (f.base_dense) NF_TRY(base_dense_launch<T>(f, (T*)c.X[0], n, 0, lq0, nullptr));    // eps -> mu + L eps, lq0 from eps
    } else if (f.base_dense && job.op == OP_ELBO) {
      NF_TRY(base_dense_launch<T>(f, (T*)c.X[0], n, 1, lq0, nullptr));                      // lq0 = log q0(x0); x0 untouched
    }
    const int last = stash ? L : 1 + ((L - 1) & 1);
    if (f.mma_mode != NF_MMA_SIMT) {
      // exact max |x0| once per chunk; every later state gets its bound from the kernel that writes it
      c.xmeta.assign(L + 1, nullptr);
      for (int i = 0; i <= L; ++i) {
        c.xmeta[i] = tc_alloc_meta(f);
        NF_REQUIRE(c.xmeta[i], "tcgen05 path: out of tensor metadata slots");
      }
      NF_TRY(tc_absmax(f, (const float*)c.X[0], n * d, c.xmeta[0]));
    }
    switch (job.op) {
      case OP_ELBO: {
        NF_REQUIRE(job.tgt, "ELBO needs a target");
        NF_TRY(sweep_forward<T>(f, c, theta, nullptr, 0, N));
        T* terms = job.terms_out ? (T*)job.terms_out + c0 : nullptr;
        {
          const size_t sm = (size_t)128 * (d + 1) * sizeof(T);
          if (sm <= 96 * 1024) {
            auto kern = elbo_head_tiled_kernel<T>;
            if (sm > 48 * 1024) NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            kern<<<(unsigned)ceil_div(n, 128), 128, sm, f.stream>>>((const T*)c.X[last], (const T*)c.X[0], (const T*)c.ld,
                                                                    job.tgt->params<T>(), base, (T)f.base_c0, d, n, (T*)c.G, terms,
                                                                    f.d_gsum + f.P, lq0);
          } else {
            elbo_head_kernel<T><<<(unsigned)ceil_div(n, 128), 128, 0, f.stream>>>(
                (const T*)c.X[last], (const T*)c.X[0], (const T*)c.ld, job.tgt->params<T>(), base, (T)f.base_c0, d, n,
                (T*)c.G, terms, f.d_gsum + f.P, lq0);
          }
        }
        NF_LAUNCH_CHECK();
        if (grad) {
          NF_TRY(alloc_backward_tmps(f, c));
          NF_TRY(sweep_backward_fwd<T>(f, c, theta));
        }
        break;
      }
      case OP_LOGLIK: {
        NF_TRY(sweep_inverse<T>(f, c, theta));
        T* terms = job.terms_out ? (T*)job.terms_out + c0 : nullptr;
        if (f.base_dense) NF_TRY(base_dense_launch<T>(f, (T*)c.X[last], n, 1, lq0, grad ? (T*)c.G : nullptr));
        loglik_head_kernel<T><<<(unsigned)ceil_div(n, 128), 128, 0, f.stream>>>(
            (const T*)c.X[last], (const T*)c.ld, base, (T)f.base_c0, d, n, (grad && !f.base_dense) ? (T*)c.G : nullptr, terms, f.d_gsum + f.P,
            lq0);
        NF_LAUNCH_CHECK();
        if (grad) {
          NF_TRY(alloc_backward_tmps(f, c));
          NF_TRY(sweep_backward_inv<T>(f, c, theta));
        }
        break;
      }
      case OP_FORWARD:
      case OP_FORWARD_STASH:
      case OP_INVERSE: {
        if (job.op == OP_INVERSE) NF_TRY(sweep_inverse<T>(f, c, theta));
        else NF_TRY(sweep_forward<T>(f, c, theta, job.bins_out, c0, N));
        if (job.y_out)
          NF_CUDA(cudaMemcpyAsync((T*)job.y_out + c0 * d, c.X[last], (size_t)n * d * sizeof(T), cudaMemcpyDeviceToDevice, f.stream));
        if (job.ld_out)
          NF_CUDA(cudaMemcpyAsync((T*)job.ld_out + c0, c.ld, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, f.stream));
        if (job.op == OP_FORWARD_STASH) {
          GenState* gs = (GenState*)f.gen_state;
          if (!gs) { gs = new GenState(); f.gen_state = gs; }
          gs->stash_chunk = c;
          gs->has_stash = true;
        }
        break;
      }
      default:
        set_error("unknown op %d", job.op);
        return NF_ERR_INVALID;
    }
  }
  return NF_OK;
}

template <typename T>
int backward_from_stash_typed(Flow& f, const void* gy_host, const void* gld_host) {
  GenState* gs = (GenState*)f.gen_state;
  NF_REQUIRE(gs && gs->has_stash, "no stashed forward pass");
  Chunk& c = gs->stash_chunk;
  const size_t es = sizeof(T);
  NF_CUDA(cudaMemsetAsync(f.d_gsum, 0, (f.P + 1) * sizeof(double), f.stream));
  NF_CUDA(cudaMemcpyAsync(c.G, gy_host, (size_t)c.n * f.dim * es, cudaMemcpyHostToDevice, f.stream));
  c.gld = f.ws_alloc((size_t)c.n * es);
  if (!c.gld) return NF_ERR_OOM;
  if (gld_host) NF_CUDA(cudaMemcpyAsync(c.gld, gld_host, (size_t)c.n * es, cudaMemcpyHostToDevice, f.stream));
  else NF_CUDA(cudaMemsetAsync(c.gld, 0, (size_t)c.n * es, f.stream));
  NF_TRY(alloc_backward_tmps(f, c));
  // the weight planes and per-tensor scales prepared by nf_forward_stash are still current
  NF_TRY(sweep_backward_fwd<T>(f, c, (const T*)f.d_theta));
  gs->has_stash = false;
  return NF_OK;
}

}  // namespace

int general_plan_workspace(Flow& f, int op, int64_t N, size_t extra_bytes) {
  if (f.all_elementwise && !(f.base_dense && (op == OP_INVERSE || op == OP_LOGLIK))) {
    f.chunk_N = N;
    return f.ws_reserve(extra_bytes + ((size_t)16 << 20) + hmc_warp_workspace_bytes(f));
  }
  const bool stash = (op == OP_ELBO || op == OP_LOGLIK || op == OP_FORWARD_STASH);
  // the plan of the previous call is reused when nothing it depends on changed (cudaMemGetInfo is a driver round trip
  // that can stall for milliseconds while the GPU is busy; a training loop asks the same question every iteration)
  if (f.plan_valid && f.plan_op == op && f.plan_N == N && f.plan_extra == extra_bytes && f.plan_limit == f.ws_limit &&
      f.plan_mode == f.mma_mode && f.ws.cap >= f.plan_cap) {
    f.chunk_N = f.plan_chunk;
    return NF_OK;
  }
  const size_t per = per_sample_bytes(f, stash);
  const size_t fixed = extra_bytes + ((size_t)32 << 20) + tc_weight_bytes(f);
  size_t budget = f.ws_limit > fixed ? f.ws_limit - fixed : 0;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
    const size_t avail = free_b + f.ws.cap;
    const size_t cap = avail > fixed + ((size_t)1 << 30) ? avail - fixed - ((size_t)1 << 30) : avail / 2;
    budget = std::min(budget, cap);
  }
  int64_t Nc = (int64_t)(budget / per);
  if (Nc >= N) Nc = N;
  else {
    Nc = (Nc / 1024) * 1024;
    NF_REQUIRE(Nc >= 1024, "workspace limit too small: %zu B per sample, budget %zu B", per, budget);
  }
  f.chunk_N = Nc;
  NF_TRY(f.ws_reserve(fixed + per * (size_t)(Nc + 1024)));
  f.plan_valid = true; f.plan_op = op; f.plan_N = N; f.plan_extra = extra_bytes; f.plan_limit = f.ws_limit;
  f.plan_mode = f.mma_mode; f.plan_chunk = Nc; f.plan_cap = fixed + per * (size_t)(Nc + 1024);
  return NF_OK;
}

int general_run(Flow& f, const GeneralJob& job) {
  if (f.all_elementwise && !(f.base_dense && (job.op == OP_INVERSE || job.op == OP_LOGLIK))) {
    set_error("operation %d is not implemented for purely elementwise (planar/radial) flows in this build", job.op);
    return NF_ERR_UNSUPPORTED;
  }
  return f.dtype == NF_F32 ? run_typed<float>(f, job) : run_typed<double>(f, job);
}

int general_backward_from_stash(Flow& f, const void* gy_host, const void* gld_host) {
  return f.dtype == NF_F32 ? backward_from_stash_typed<float>(f, gy_host, gld_host)
                           : backward_from_stash_typed<double>(f, gy_host, gld_host);
}

void general_release(Flow& f) {
  delete (GenState*)f.gen_state;
  f.gen_state = nullptr;
  tc_release(f);
}

int base_sample_dev(Flow& f, int64_t N, uint64_t seed, void* z_dev) {
  const int d = f.dim;
  const bool plain = f.base_is_standard || f.base_dense;
  if (f.dtype == NF_F32)
    base_sample_kernel<float><<<base_sample_grid(N, d, 0), 256, 0, f.stream>>>((float*)z_dev, plain ? nullptr : (const float*)f.d_base, d, N, seed, 0);
  else
    base_sample_kernel<double><<<base_sample_grid(N, d, 0), 256, 0, f.stream>>>((double*)z_dev, plain ? nullptr : (const double*)f.d_base, d, N, seed, 0);
  NF_LAUNCH_CHECK();
  if (f.base_dense) {       // unwhiten!(Sigma, x) .+ mu of reference ext/NormalizingFlowsCUDAExt.jl:43-47
    if (f.dtype == NF_F32) NF_TRY(base_dense_launch<float>(f, (float*)z_dev, N, 0, nullptr, nullptr));
    else NF_TRY(base_dense_launch<double>(f, (double*)z_dev, N, 0, nullptr, nullptr));
  }
  return NF_OK;
}

int rqs_bin_search_host(int dtype, const void* knots_host, const void* v_host, int64_t M, int K, int32_t* bins_out) {
  const size_t es = dtype == NF_F64 ? 8 : 4;
  void *dk = nullptr, *dv = nullptr;
  int32_t* db = nullptr;
  NF_CUDA(cudaMalloc(&dk, (size_t)M * (K + 1) * es));
  NF_CUDA(cudaMalloc(&dv, (size_t)M * es));
  NF_CUDA(cudaMalloc((void**)&db, (size_t)M * sizeof(int32_t)));
  NF_CUDA(cudaMemcpy(dk, knots_host, (size_t)M * (K + 1) * es, cudaMemcpyHostToDevice));
  NF_CUDA(cudaMemcpy(dv, v_host, (size_t)M * es, cudaMemcpyHostToDevice));
  if (dtype == NF_F64) rqs_bin_search_kernel<double><<<(unsigned)ceil_div(M, 256), 256>>>((const double*)dk, (const double*)dv, M, K, db);
  else rqs_bin_search_kernel<float><<<(unsigned)ceil_div(M, 256), 256>>>((const float*)dk, (const float*)dv, M, K, db);
  NF_LAUNCH_CHECK();
  NF_CUDA(cudaMemcpy(bins_out, db, (size_t)M * sizeof(int32_t), cudaMemcpyDeviceToHost));
  cudaFree(dk); cudaFree(dv); cudaFree(db);
  return NF_OK;
}

}  // namespace nf
