// Host-side flow / target descriptors of libnfcuda.
//
// A Flow mirrors `transformed(q0, reduce(∘, Ls))` (reference src/flows/utils.jl:23-26): `layers` is Ls
// in theta (= Optimisers.destructure, reference src/NormalizingFlows.jl:67) order; the transform
// applies layers.back() first.
#pragma once
#include "common.cuh"
#include "targets.cuh"
#include <algorithm>
#include <memory>
#include <map>
#include <string>

namespace nf {

struct EwLayerMeta {
  int kind;
  int aux;              // leapfrog: number of steps
  int64_t theta_off;
};

// fnn(in, hdims, out; output_activation) of reference src/flows/utils.jl:71-100.
// dims = [in, h1, ..., out]; Dense i has W (dims[i+1] x dims[i], column major in theta == row-major
// [dims[i]][dims[i+1]]) at w_off[i] and bias at b_off[i].
struct MLPDesc {
  std::vector<int> dims;
  std::vector<int64_t> w_off, b_off;   // absolute offsets into theta
  int out_act = 0;                     // 0 none, 1 tanh
  int n_dense() const { return (int)dims.size() - 1; }
  int64_t n_params() const {
    int64_t n = 0;
    for (int i = 0; i + 1 < (int)dims.size(); ++i) n += (int64_t)dims[i] * dims[i + 1] + dims[i + 1];
    return n;
  }
};

struct LayerDesc {
  int kind = 0;
  int64_t theta_off = 0, n_params = 0;
  // couplings (PartitionMask, App. A.3): idx1 = transformed coordinates, idx2 = sorted complement
  std::vector<int> idx1, idx2;
  int* d_idx1 = nullptr;
  int* d_idx2 = nullptr;
  int* d_pos = nullptr;                // [dim]: position of a column inside idx1, or -1
  std::vector<MLPDesc> mlps;           // affine: {s, t}; spline: {nn}
  int K = 0;
  double B = 0;
};

// Optional per-kernel-class timing with CUDA events on the flow's stream (bench.py roofline numbers).
struct Profiler {
  bool on = false;
  struct Rec { cudaEvent_t a, b; };
  std::map<std::string, std::vector<Rec>> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void begin(const std::string& key, cudaStream_t st) {
    if (!on) return;
    Rec r{get(), get()};
    cudaEventRecord(r.a, st);
    recs[key].push_back(r);
    cur = key;
  }
  void end(cudaStream_t st) {
    if (!on || cur.empty()) return;
    cudaEventRecord(recs[cur].back().b, st);
    cur.clear();
  }
  // call after the stream has been synchronised; returns launches and total ms, recycles the events
  void collect(const std::string& key, int64_t* count, double* ms) {
    *count = 0; *ms = 0;
    auto it = recs.find(key);
    if (it == recs.end()) return;
    for (auto& r : it->second) {
      float t = 0; cudaEventElapsedTime(&t, r.a, r.b);
      *ms += t; ++*count;
      pool.push_back(r.a); pool.push_back(r.b);
    }
    recs.erase(it);
  }
  std::string keys() const { std::string k; for (auto& kv : recs) { if (!k.empty()) k += ","; k += kv.first; } return k; }
  std::string cur;
};

struct Target;
struct Workspace {
  char* base = nullptr;
  size_t cap = 0, off = 0;
};

struct Flow {
  int dim = 0, dtype = NF_F32, device = 0;
  std::vector<LayerDesc> layers;
  int64_t P = 0;
  bool all_elementwise = true, any_elementwise = false;
  bool hamiltonian = false;            // has NF_MOMENTUM_AFFINE / NF_LEAPFROG layers (z = [x, rho])
  struct Target* score_target = nullptr;   // owned copy of the LeapFrog layers' target
  int mma_mode = NF_MMA_SIMT;
  size_t ws_limit = (size_t)64 << 30;
  // cached workspace plan (general_plan_workspace)
  bool plan_valid = false;
  int plan_op = 0, plan_mode = 0;
  int64_t plan_N = 0, plan_chunk = 0;
  size_t plan_extra = 0, plan_limit = 0, plan_cap = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // host-supplied inputs of large batches arrive in two halves: the second half's copy (copy_stream) overlaps the first half's
  // compute; the chunk loop of the layered path waits on in_ev before it touches rows >= in_split  (stage_host_input)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t in_ev = nullptr;
  bool in_ev_armed = false;
  double last_ms = 0;

  // base distribution q0 = MvNormal(mu, Diagonal(sigma^2))
  bool base_is_standard = true;
  std::vector<double> base_mu, base_sigma;
  void* d_base = nullptr;              // dtype[2*dim]: mu then sigma
  double base_c0 = 0;
  // full covariance: q0 = MvNormal(mu, L L^T) (nf_flow_set_base_chol); d_base_L = dtype[dim*dim + dim]: L (row major, lower) then mu
  bool base_dense = false;
  void* d_base_L = nullptr;

  // elementwise-path metadata
  EwLayerMeta* d_ew_meta = nullptr;
  int* d_ew_kinds = nullptr;

  // persistent device buffers
  void* d_theta = nullptr;             // dtype[P]
  double* d_gsum = nullptr;            // double[P+1] un-normalised gradient sums + ELBO sum
  void* d_out = nullptr;               // dtype[P+1] scaled outputs
  void* d_adam = nullptr;              // dtype[2P]: Adam first / second moments (on-device optimiser loop)
  double* d_stats = nullptr;           // [iters][2]: loss, |g|^2
  int64_t* d_iter = nullptr;           // device iteration counter (+ 2 uints) of the graph-replayed training loop
  int stats_cap = 0;
  void* h_pinned = nullptr;            // pinned staging (P+1 of dtype + 1 double)
  size_t h_pinned_bytes = 0;

  Workspace ws;
  Profiler prof;
  int64_t stash_N = 0;                 // samples of the forward pass kept by nf_forward_stash for nf_backward (0: none)
  // every entry point that recycles the workspace also drops a stashed forward pass (its pointers live in the workspace),
  // so a later nf_backward fails with 'needs a preceding nf_forward_stash' instead of reading recycled memory
  void ws_reset() { ws.off = 0; stash_N = 0; }
  void* ws_alloc(size_t bytes);        // bump allocation, 256 B aligned; nullptr (+error) when exhausted
  int ws_reserve(size_t bytes);        // make sure capacity >= bytes (reallocates; invalidates pointers)

  // layered path: chunk size chosen by general_plan_workspace, opaque per-flow state
  int64_t chunk_N = 0;
  // device Philox draws: row of the global batch this flow's first draw corresponds to (data-parallel shards draw the
  // rows [offset, offset + N) of ONE global draw matrix, so G devices reproduce the single-device batch)
  int64_t draw_row_offset = 0;
  void* gen_state = nullptr;
  void* tc_state = nullptr;            // tcgen05 path: prepared weight planes, tensor-map cache

  size_t esize() const { return dtype == NF_F64 ? 8 : 4; }
};

struct Target {
  int kind = 0, dim = 0;
  bool joint = false;        // dim is then 2 * (inner dim)
  int n_data = 0;            // LogReg: observations; d_vec holds X[n][dim] then y[n]
  std::vector<double> p;
  double c0 = 0;
  void* d_vec_f32 = nullptr;
  void* d_vec_f64 = nullptr;
  template <typename T> TargetParams<T> params() const {
    TargetParams<T> tp;
    tp.kind = kind; tp.dim = dim;
    tp.p0 = p.size() > 0 ? (T)p[0] : T(0);
    tp.p1 = p.size() > 1 ? (T)p[1] : T(0);
    tp.c0 = (T)c0;
    tp.vec = (const T*)(sizeof(T) == 4 ? d_vec_f32 : d_vec_f64);
    tp.n_data = n_data;
    tp.joint = joint ? 1 : 0;
    if (joint) tp.dim = dim / 2;
    return tp;
  }
};

// hmc_warp.cu: warp-per-sample Hamiltonian flows (dim = 2h, h <= 128; forward direction)
bool hmc_warp_qualifies(const Flow& f, const struct Target* tgt);
size_t hmc_warp_workspace_bytes(const Flow& f);
template <typename T>
int hmc_warp_run(Flow& f, const struct Target* tgt, const void* theta_dev, int64_t N, const void* z0_dev, uint64_t seed, bool want_grad,
                 void* y_out, void* ld_out, void* terms_out, double* gsum_dev);
template <typename T>
int hmc_warp_inverse(Flow& f, const void* theta_dev, int64_t N, const void* y_dev, bool head, bool want_grad, void* x_out, void* ld_out,
                     void* terms_out, double* gsum_dev);
extern template int hmc_warp_inverse<float>(Flow&, const void*, int64_t, const void*, bool, bool, void*, void*, void*, double*);
extern template int hmc_warp_inverse<double>(Flow&, const void*, int64_t, const void*, bool, bool, void*, void*, void*, double*);
extern int g_opt_rqs_planes;    // nf_set_option("rqs_planes", v): spline backward writes the conditioner-output gradient as split planes (1, default),
                                // as an fp32 matrix plus a split pass (0), or as planes with a deliberately wrong predicted scale so that the
                                // redo pass runs (2: tests)
extern int g_opt_hmc_warp;      // nf_set_option("hmc_warp", 1): route every qualifying Hamiltonian flow through hmc_warp.cu (tests)
extern template int hmc_warp_run<float>(Flow&, const struct Target*, const void*, int64_t, const void*, uint64_t, bool, void*, void*, void*, double*);
extern template int hmc_warp_run<double>(Flow&, const struct Target*, const void*, int64_t, const void*, uint64_t, bool, void*, void*, void*, double*);

// elementwise_{fwd,inv}_{f32,f64}.cu
template <typename T, bool INV>
int ew_run_dir(Flow& f, const Target* tgt, const void* theta_dev, int64_t N, const void* z0_dev, uint64_t seed,
               bool want_grad, void* y_out, void* ld_out, void* terms_out, double* gsum_dev, bool head);
extern template int ew_run_dir<float, false>(Flow&, const Target*, const void*, int64_t, const void*, uint64_t, bool, void*, void*, void*, double*, bool);
extern template int ew_run_dir<float, true>(Flow&, const Target*, const void*, int64_t, const void*, uint64_t, bool, void*, void*, void*, double*, bool);
extern template int ew_run_dir<double, false>(Flow&, const Target*, const void*, int64_t, const void*, uint64_t, bool, void*, void*, void*, double*, bool);
extern template int ew_run_dir<double, true>(Flow&, const Target*, const void*, int64_t, const void*, uint64_t, bool, void*, void*, void*, double*, bool);
template <typename T>
inline int ew_run(Flow& f, const Target* tgt, const void* theta_dev, int64_t N, const void* z0_dev, uint64_t seed,
                  bool want_grad, void* y_out, void* ld_out, void* terms_out, double* gsum_dev, bool inverse = false,
                  bool head = false) {
  return inverse ? ew_run_dir<T, true>(f, tgt, theta_dev, N, z0_dev, seed, want_grad, y_out, ld_out, terms_out, gsum_dev, head)
                 : ew_run_dir<T, false>(f, tgt, theta_dev, N, z0_dev, seed, want_grad, y_out, ld_out, terms_out, gsum_dev, head);
}

// a run of planar / radial / shift layers inside a layered (coupling) flow, in the forward or the inverse direction:
// forward pass (Xout, ld +=) or backward pass (G in place, gsum +=)
template <typename T, bool INV>
int ew_segment_dir(Flow& f, int l0, int Lseg, const void* theta_dev, int64_t N, const void* Xin, void* Xout, void* ld, bool backward,
                   void* G, const void* gld, double* gsum);
extern template int ew_segment_dir<float, false>(Flow&, int, int, const void*, int64_t, const void*, void*, void*, bool, void*, const void*, double*);
extern template int ew_segment_dir<float, true>(Flow&, int, int, const void*, int64_t, const void*, void*, void*, bool, void*, const void*, double*);
extern template int ew_segment_dir<double, false>(Flow&, int, int, const void*, int64_t, const void*, void*, void*, bool, void*, const void*, double*);
extern template int ew_segment_dir<double, true>(Flow&, int, int, const void*, int64_t, const void*, void*, void*, bool, void*, const void*, double*);
template <typename T>
inline int ew_segment(Flow& f, int l0, int Lseg, const void* theta_dev, int64_t N, const void* Xin, void* Xout, void* ld, bool backward,
                      void* G, const void* gld, double* gsum, bool inverse = false) {
  return inverse ? ew_segment_dir<T, true>(f, l0, Lseg, theta_dev, N, Xin, Xout, ld, backward, G, gld, gsum)
                 : ew_segment_dir<T, false>(f, l0, Lseg, theta_dev, N, Xin, Xout, ld, backward, G, gld, gsum);
}

// single-launch Adam training loop for small batches of elementwise flows (elementwise.cu: ew_train_kernel)
template <typename T>
int ew_train(Flow& f, const Target* tgt, int64_t N, uint64_t seed, int n_iters, int t0, double eta, double b1, double b2,
             double eps, void* m_dev, void* v_dev);

}  // namespace nf
