// libnfcuda C ABI (include/nfcuda.h): handle management, host<->device staging, dispatch.
#include "flow.hpp"
#include "general.hpp"
#include "tc_gemm.hpp"
#include <chrono>
#include <cstring>
#include <mutex>
#include <dlfcn.h>
#include <nccl.h>

namespace nf {

static thread_local std::string g_last_error;
thread_local int64_t g_launch_count = 0;
int g_opt_hmc_warp = getenv("NFCUDA_HMC_WARP") ? atoi(getenv("NFCUDA_HMC_WARP")) : 0;
// 0: two-team streaming kernel (fused_coupling.cuh, default), 1: 128-column-MMA kernel (fused_coupling_w128.cuh)
int g_opt_fused_variant = getenv("NFCUDA_FUSED_VARIANT") ? atoi(getenv("NFCUDA_FUSED_VARIANT")) : 0;
int g_opt_fused_coupling = !(getenv("NFCUDA_FUSED") && atoi(getenv("NFCUDA_FUSED")) == 0);
int g_opt_rqs_planes = getenv("NFCUDA_RQS_PLANES") ? atoi(getenv("NFCUDA_RQS_PLANES")) : 1;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

void* Flow::ws_alloc(size_t bytes) {
  const size_t a = (ws.off + 255) & ~(size_t)255;
  if (a + bytes > ws.cap) {
    set_error("workspace exhausted: need %zu B at offset %zu, capacity %zu B", bytes, a, ws.cap);
    return nullptr;
  }
  ws.off = a + bytes;
  return ws.base + a;
}

int Flow::ws_reserve(size_t bytes) {
  if (bytes <= ws.cap) return NF_OK;
  if (ws.base) { NF_CUDA(cudaStreamSynchronize(stream)); NF_CUDA(cudaFree(ws.base)); ws.base = nullptr; ws.cap = 0; }
  bytes = (size_t)round_up((int64_t)bytes, 1 << 20);
  NF_CUDA(cudaMalloc((void**)&ws.base, bytes));
  ws.cap = bytes;
  return NF_OK;
}

template <typename T>
__global__ void target_logp_kernel(TargetParams<T> tp, const T* __restrict__ X, int64_t N, T* __restrict__ lp, T* __restrict__ G) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  const int d = tp.joint ? 2 * tp.dim : tp.dim;
  lp[r] = target_logp_score<T, 0>(tp, X + r * d, G + r * d);
}

template <typename T>
__global__ void scale_out_kernel(const double* __restrict__ gsum, int64_t n, double factor, T* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (T)(gsum[i] * factor);
}

// Optimisers.Adam step on the device (reference src/optimize.jl:99), in the flow's element type:
//   m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ; theta <- theta - eta * (m / (1-b1^t)) / (sqrt(v / (1-b2^t)) + eps)
// g = -gsum/N is the gradient of the loss -vo.  Also records (loss, |g|^2) of this iteration.
template <typename T>
__global__ void adam_step_kernel(const double* __restrict__ gsum, int64_t P, double inv_n, T b1, T b2, T omb1t, T omb2t, T eta,
                                 T eps, T* __restrict__ theta, T* __restrict__ m, T* __restrict__ v, double* __restrict__ stat) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double g2 = 0;
  if (i < P) {
    const T g = (T)(-gsum[i] * inv_n);
    const T mi = b1 * m[i] + (T(1) - b1) * g;
    const T vi = b2 * v[i] + (T(1) - b2) * g * g;
    m[i] = mi; v[i] = vi;
    theta[i] -= mi / omb1t / (Num<T>::sqrt(vi / omb2t) + eps) * eta;
    g2 = (double)g * (double)g;
  }
  g2 = warp_sum(g2);
  if ((threadIdx.x & 31) == 0 && g2 != 0) atomicAdd(&stat[1], g2);
  if (i == 0) stat[0] = -gsum[P] * inv_n;
}

// Same step for the CUDA-graph-replayed loop: the iteration index lives on the device (one captured launch serves every replay);
// the last thread block to finish advances it.
template <typename T>
__global__ void adam_step_graph_kernel(const double* __restrict__ gsum, int64_t P, double inv_n, double b1d, double b2d, int t0, T eta,
                                       T eps, T* __restrict__ theta, T* __restrict__ m, T* __restrict__ v, double* __restrict__ stats,
                                       int64_t* __restrict__ iter, unsigned int* __restrict__ done) {
  const int64_t it = *iter;
  const int step = t0 + (int)it + 1;
  const T b1 = (T)b1d, b2 = (T)b2d;
  const T omb1t = (T)(1.0 - pow(b1d, (double)step)), omb2t = (T)(1.0 - pow(b2d, (double)step));
  double* stat = stats + 2 * it;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double g2 = 0;
  if (i < P) {
    const T g = (T)(-gsum[i] * inv_n);
    const T mi = b1 * m[i] + (T(1) - b1) * g;
    const T vi = b2 * v[i] + (T(1) - b2) * g * g;
    m[i] = mi; v[i] = vi;
    theta[i] -= mi / omb1t / (Num<T>::sqrt(vi / omb2t) + eps) * eta;
    g2 = (double)g * (double)g;
  }
  g2 = warp_sum(g2);
  if ((threadIdx.x & 31) == 0 && g2 != 0) atomicAdd(&stat[1], g2);
  if (i == 0) stat[0] = -gsum[P] * inv_n;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(done, 1u) == gridDim.x - 1) { *done = 0; *iter = it + 1; }   // every block has read `it` before this point
  }
}

static int check_device() {
  int dev = 0;
  NF_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  NF_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("libnfcuda is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor);
    return NF_ERR_CUDA;
  }
  return NF_OK;
}

static int clone_target(const Target& src, Target& dst) {
  dst = src;
  dst.d_vec_f32 = dst.d_vec_f64 = nullptr;
  if (src.d_vec_f32) {
    const int inner_dim = src.joint ? src.dim / 2 : src.dim;
    const int n = src.kind == NF_TARGET_LOGREG ? src.n_data * inner_dim + src.n_data : 2 * inner_dim;
    NF_CUDA(cudaMalloc(&dst.d_vec_f32, n * sizeof(float)));
    NF_CUDA(cudaMalloc(&dst.d_vec_f64, n * sizeof(double)));
    NF_CUDA(cudaMemcpy(dst.d_vec_f32, src.d_vec_f32, n * sizeof(float), cudaMemcpyDeviceToDevice));
    NF_CUDA(cudaMemcpy(dst.d_vec_f64, src.d_vec_f64, n * sizeof(double), cudaMemcpyDeviceToDevice));
  }
  return NF_OK;
}

static int build_mlp(MLPDesc& m, int n_in, const int* hdims, int n_hidden, int n_out, int out_act, int64_t& off) {
  m.dims.clear();
  m.dims.push_back(n_in);
  for (int i = 0; i < n_hidden; ++i) {
    NF_REQUIRE(hdims[i] > 0, "hidden width must be positive");
    m.dims.push_back(hdims[i]);
  }
  m.dims.push_back(n_out);
  m.out_act = out_act;
  for (int i = 0; i + 1 < (int)m.dims.size(); ++i) {
    m.w_off.push_back(off); off += (int64_t)m.dims[i] * m.dims[i + 1];
    m.b_off.push_back(off); off += m.dims[i + 1];
  }
  return NF_OK;
}

static int upload_base(Flow& f) {
  const int d = f.dim;
  f.base_is_standard = true;
  double c0 = -0.5 * d * NF_LOG2PI;
  for (int k = 0; k < d; ++k) {
    if (f.base_mu[k] != 0.0 || f.base_sigma[k] != 1.0) f.base_is_standard = false;
    c0 -= std::log(f.base_sigma[k]);
  }
  f.base_c0 = c0;
  if (!f.d_base) NF_CUDA(cudaMalloc(&f.d_base, 2 * d * sizeof(double)));
  if (f.dtype == NF_F32) {
    std::vector<float> h(2 * d);
    for (int k = 0; k < d; ++k) { h[k] = (float)f.base_mu[k]; h[d + k] = (float)f.base_sigma[k]; }
    NF_CUDA(cudaMemcpy(f.d_base, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  } else {
    std::vector<double> h(2 * d);
    for (int k = 0; k < d; ++k) { h[k] = f.base_mu[k]; h[d + k] = f.base_sigma[k]; }
    NF_CUDA(cudaMemcpy(f.d_base, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  return NF_OK;
}

static void flow_destroy(Flow* f);

static int flow_create(nf_flow_t* out, const nf_layer_desc* descs, int n_layers, int dim, int dtype) {
  NF_REQUIRE(out && descs, "null argument");
  NF_REQUIRE(n_layers > 0 && dim > 0, "need at least one layer and dim > 0");
  NF_REQUIRE(dtype == NF_F32 || dtype == NF_F64, "dtype must be NF_F32 or NF_F64");
  NF_TRY(check_device());
  // flow_destroy (not plain delete) on every failure path: it also frees the device buffers, stream and events created so far
  std::unique_ptr<Flow, void (*)(Flow*)> f(new Flow(), flow_destroy);
  f->dim = dim; f->dtype = dtype;
  NF_CUDA(cudaGetDevice(&f->device));
  int64_t off = 0;
  for (int i = 0; i < n_layers; ++i) {
    const nf_layer_desc& ds = descs[i];
    LayerDesc L;
    L.kind = ds.kind; L.theta_off = off;
    switch (ds.kind) {
      case NF_PLANAR: off += 2 * dim + 1; f->any_elementwise = true; break;
      case NF_RADIAL: off += dim + 2; f->any_elementwise = true; break;
      case NF_SHIFT: case NF_SCALE: off += dim; f->any_elementwise = true; break;
      case NF_MOMENTUM_AFFINE: case NF_LEAPFROG: {
        // h a power of two <= 32: fused one-thread-per-sample kernel (both directions); any h <= 128: warp-per-sample kernel
        // (hmc_warp.cu, forward direction: ELBO / gradient / transform)
        NF_REQUIRE(dim >= 2 && (dim & 1) == 0 && dim <= 256, "layer %d: Hamiltonian layers need dim = 2h with h <= 128, got %d", i, dim);
        f->any_elementwise = true; f->hamiltonian = true;
        if (ds.kind == NF_MOMENTUM_AFFINE) { off += dim; break; }
        NF_REQUIRE(ds.n_steps >= 1, "layer %d: leapfrog needs n_steps >= 1", i);
        const Target* st = reinterpret_cast<const Target*>(ds.score_target);
        NF_REQUIRE(st && !st->joint && st->dim == dim / 2, "layer %d: leapfrog needs a score target over dim/2 = %d coordinates", i, dim / 2);
        NF_REQUIRE(target_has_hvp(st->kind), "layer %d: leapfrog supports Banana, Funnel and DiagNormal score targets", i);
        if (!f->score_target) {
          f->score_target = new Target();
          NF_TRY(clone_target(*st, *f->score_target));
        } else {
          NF_REQUIRE(f->score_target->kind == st->kind && f->score_target->p == st->p && f->score_target->n_data == st->n_data,
                     "layer %d: all leapfrog layers of a flow must share one score target", i);
        }
        L.K = ds.n_steps;
        off += dim / 2;
        break;
      }
      case NF_AFFINE_COUPLING: case NF_SPLINE_COUPLING: {
        f->all_elementwise = false;
        NF_REQUIRE(ds.mask_idx && ds.n_mask > 0 && ds.n_mask < dim, "layer %d: coupling needs 0 < n_mask < dim", i);
        NF_REQUIRE(ds.hdims && ds.n_hidden > 0, "layer %d: coupling needs hidden widths", i);
        std::vector<char> used(dim, 0);
        for (int k = 0; k < ds.n_mask; ++k) {
          NF_REQUIRE(ds.mask_idx[k] >= 0 && ds.mask_idx[k] < dim && !used[ds.mask_idx[k]], "layer %d: bad mask index", i);
          used[ds.mask_idx[k]] = 1;
          L.idx1.push_back(ds.mask_idx[k]);
        }
        for (int k = 0; k < dim; ++k) if (!used[k]) L.idx2.push_back(k);
        const int c = ds.n_mask, cbar = dim - c;
        if (ds.kind == NF_AFFINE_COUPLING) {
          L.mlps.resize(2);
          NF_TRY(build_mlp(L.mlps[0], cbar, ds.hdims, ds.n_hidden, c, 1, off));   // s: ... tanh (realnvp.jl:50)
          NF_TRY(build_mlp(L.mlps[1], cbar, ds.hdims, ds.n_hidden, c, 0, off));   // t          (realnvp.jl:51)
        } else {
          NF_REQUIRE(ds.K >= 2 && ds.B > 0, "layer %d: spline needs K >= 2 and B > 0", i);
          NF_REQUIRE(ds.K <= 64, "layer %d: K <= 64 supported", i);
          L.K = ds.K; L.B = ds.B;
          L.mlps.resize(1);
          NF_TRY(build_mlp(L.mlps[0], cbar, ds.hdims, ds.n_hidden, (3 * ds.K - 1) * c, 0, off));  // neuralspline.jl:55-56
        }
        break;
      }
      default:
        set_error("layer %d: unknown kind %d", i, ds.kind);
        return NF_ERR_INVALID;
    }
    L.n_params = off - L.theta_off;
    f->layers.push_back(std::move(L));
    if (ds.kind == NF_AFFINE_COUPLING || ds.kind == NF_SPLINE_COUPLING) {
      // device copies of the masks are created on the layer the flow already owns, so a failure below frees them
      LayerDesc& Lo = f->layers.back();
      const int c = (int)Lo.idx1.size(), cbar = (int)Lo.idx2.size();
      NF_CUDA(cudaMalloc((void**)&Lo.d_idx1, c * sizeof(int)));
      NF_CUDA(cudaMalloc((void**)&Lo.d_idx2, cbar * sizeof(int)));
      NF_CUDA(cudaMemcpy(Lo.d_idx1, Lo.idx1.data(), c * sizeof(int), cudaMemcpyHostToDevice));
      NF_CUDA(cudaMemcpy(Lo.d_idx2, Lo.idx2.data(), cbar * sizeof(int), cudaMemcpyHostToDevice));
      std::vector<int> pos(dim, -1);
      for (int k = 0; k < c; ++k) pos[Lo.idx1[k]] = k;
      NF_CUDA(cudaMalloc((void**)&Lo.d_pos, dim * sizeof(int)));
      NF_CUDA(cudaMemcpy(Lo.d_pos, pos.data(), dim * sizeof(int), cudaMemcpyHostToDevice));
    }
  }
  f->P = off;
  f->mma_mode = (dtype == NF_F32) ? NF_MMA_F16X3 : NF_MMA_SIMT;
  if (const char* e = getenv("NFCUDA_MMA")) {
    if (!strcmp(e, "simt")) f->mma_mode = NF_MMA_SIMT;
    else if (!strcmp(e, "f16x1") && dtype == NF_F32) f->mma_mode = NF_MMA_F16X1;
    else if (!strcmp(e, "f16x3") && dtype == NF_F32) f->mma_mode = NF_MMA_F16X3;
  }
  NF_CUDA(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
  NF_CUDA(cudaEventCreate(&f->ev0));
  NF_CUDA(cudaEventCreate(&f->ev1));
  f->base_mu.assign(dim, 0.0);
  f->base_sigma.assign(dim, 1.0);
  NF_TRY(upload_base(*f));
  {
    std::vector<EwLayerMeta> meta(n_layers);
    std::vector<int> kinds(n_layers);
    for (int i = 0; i < n_layers; ++i) { meta[i].kind = f->layers[i].kind; meta[i].aux = f->layers[i].K; meta[i].theta_off = f->layers[i].theta_off; kinds[i] = f->layers[i].kind; }
    NF_CUDA(cudaMalloc((void**)&f->d_ew_meta, n_layers * sizeof(EwLayerMeta)));
    NF_CUDA(cudaMalloc((void**)&f->d_ew_kinds, n_layers * sizeof(int)));
    NF_CUDA(cudaMemcpy(f->d_ew_meta, meta.data(), n_layers * sizeof(EwLayerMeta), cudaMemcpyHostToDevice));
    NF_CUDA(cudaMemcpy(f->d_ew_kinds, kinds.data(), n_layers * sizeof(int), cudaMemcpyHostToDevice));
  }
  const size_t es = f->esize();
  NF_CUDA(cudaMalloc(&f->d_theta, (f->P + 1) * es));
  NF_CUDA(cudaMalloc((void**)&f->d_gsum, (f->P + 1) * sizeof(double)));
  NF_CUDA(cudaMalloc(&f->d_out, (f->P + 1) * es));
  f->h_pinned_bytes = (((f->P + 1) * es + 15) / 16) * 16 + 64;
  NF_CUDA(cudaMallocHost(&f->h_pinned, f->h_pinned_bytes));
  NF_TRY(f->ws_reserve((size_t)8 << 20));
  *out = reinterpret_cast<nf_flow_t>(f.release());
  return NF_OK;
}

static void flow_destroy(Flow* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  if (f->stream) cudaStreamSynchronize(f->stream);
  general_release(*f);
  for (auto& L : f->layers) { cudaFree(L.d_idx1); cudaFree(L.d_idx2); cudaFree(L.d_pos); }
  cudaFree(f->d_base); cudaFree(f->d_base_L); cudaFree(f->d_ew_meta); cudaFree(f->d_ew_kinds);
  cudaFree(f->d_theta); cudaFree(f->d_gsum); cudaFree(f->d_out); cudaFreeHost(f->h_pinned);
  cudaFree(f->d_adam); cudaFree(f->d_stats); cudaFree(f->d_iter);
  cudaFree(f->ws.base);
  if (f->score_target) { cudaFree(f->score_target->d_vec_f32); cudaFree(f->score_target->d_vec_f64); delete f->score_target; }
  if (f->ev0) cudaEventDestroy(f->ev0);
  if (f->ev1) cudaEventDestroy(f->ev1);
  if (f->copy_stream) cudaStreamDestroy(f->copy_stream);
  if (f->in_ev) cudaEventDestroy(f->in_ev);
  if (f->stream) cudaStreamDestroy(f->stream);
  delete f;
}

// One objective / transform evaluation on device buffers.
struct Job {
  int op = OP_ELBO;
  const Target* tgt = nullptr;
  const void* theta_dev = nullptr;
  const void* in_dev = nullptr;       // z0 / x / y; nullptr -> device Philox draws (OP_ELBO / OP_FORWARD)
  int64_t N = 0;
  uint64_t seed = 0;
  const int64_t* seed_iter_dev = nullptr;
  bool want_grad = false;
  void* y_out = nullptr;              // [N, d]
  void* ld_out = nullptr;             // [N]
  void* terms_out = nullptr;          // [N]
  int32_t* bins_out = nullptr;
};

// Runs the job; leaves un-normalised sums in f.d_gsum (gradient sums [0,P), objective sum at [P]).
static int run_job(Flow& f, const Job& j) {
  NF_REQUIRE(j.N > 0, "N must be positive");
  // (a full-covariance base in the inverse direction goes through the layered path: the whole flow is one elementwise segment there)
  const bool dense_inv = f.base_dense && (j.op == OP_INVERSE || j.op == OP_LOGLIK);
  if (f.all_elementwise && !dense_inv && (j.op == OP_ELBO || j.op == OP_FORWARD || j.op == OP_INVERSE || j.op == OP_LOGLIK)) {
    const bool inv = j.op == OP_INVERSE || j.op == OP_LOGLIK;
    const bool head = j.op == OP_LOGLIK;
    double* gs = (j.op == OP_ELBO || j.op == OP_LOGLIK) ? f.d_gsum : nullptr;
    NF_REQUIRE(!inv || j.in_dev, "inverse / log-likelihood need input samples");
    if (f.dtype == NF_F32)
      return ew_run<float>(f, j.tgt, j.theta_dev, j.N, j.in_dev, j.seed, j.want_grad, j.y_out, j.ld_out, j.terms_out, gs, inv, head);
    return ew_run<double>(f, j.tgt, j.theta_dev, j.N, j.in_dev, j.seed, j.want_grad, j.y_out, j.ld_out, j.terms_out, gs, inv, head);
  }
  GeneralJob g;
  g.op = j.op; g.tgt = j.tgt; g.theta_dev = j.theta_dev; g.in_dev = j.in_dev; g.N = j.N; g.seed = j.seed; g.seed_iter_dev = j.seed_iter_dev;
  g.want_grad = j.want_grad; g.y_out = j.y_out; g.ld_out = j.ld_out; g.terms_out = j.terms_out; g.bins_out = j.bins_out;
  return general_run(f, g);
}

static int scale_outputs(Flow& f, double factor, void* out_dev, int64_t n) {
  const int threads = 256;
  const int blocks = (int)ceil_div(n, threads);
  if (f.dtype == NF_F32) scale_out_kernel<float><<<blocks, threads, 0, f.stream>>>(f.d_gsum, n, factor, (float*)out_dev);
  else scale_out_kernel<double><<<blocks, threads, 0, f.stream>>>(f.d_gsum, n, factor, (double*)out_dev);
  NF_LAUNCH_CHECK();
  return NF_OK;
}

// value+grad with everything already on the device
// NFCUDA_TRACE_HOST=<ms>: report calls whose host wall time exceeds the device time by more than <ms> (where the host
// thread was held up: planning, enqueue, or the final wait).
struct HostTrace {
  double thr = -1;
  std::chrono::steady_clock::time_point t[5];
  HostTrace() { if (const char* e = getenv("NFCUDA_TRACE_HOST")) thr = atof(e); }
  void mark(int i) { if (thr >= 0) t[i] = std::chrono::steady_clock::now(); }
  double ms(int a, int b) const { return std::chrono::duration<double, std::milli>(t[b] - t[a]).count(); }
};
static HostTrace g_trace;

static int value_and_grad_dev(Flow& f, int op, const Target* tgt, const void* theta_dev, int64_t N, const void* in_dev,
                              uint64_t seed, double scale, double* value_out, void* grad_dev_out) {
  Job j;
  j.op = op; j.tgt = tgt; j.theta_dev = theta_dev; j.in_dev = in_dev; j.N = N; j.seed = seed;
  j.want_grad = grad_dev_out != nullptr;
  g_trace.mark(1);
  NF_CUDA(cudaEventRecord(f.ev0, f.stream));
  NF_TRY(run_job(f, j));
  const double factor = scale / (double)N;
  if (grad_dev_out) NF_TRY(scale_outputs(f, factor, grad_dev_out, f.P));
  NF_CUDA(cudaEventRecord(f.ev1, f.stream));
  // the objective sum comes back through the pinned staging buffer (a pageable destination would make the copy synchronous)
  double* vsum = reinterpret_cast<double*>((char*)f.h_pinned + f.h_pinned_bytes - 16);
  NF_CUDA(cudaMemcpyAsync(vsum, f.d_gsum + f.P, sizeof(double), cudaMemcpyDeviceToHost, f.stream));
  g_trace.mark(2);
  NF_CUDA(cudaStreamSynchronize(f.stream));
  g_trace.mark(3);
  float ms = 0;
  cudaEventElapsedTime(&ms, f.ev0, f.ev1);
  f.last_ms = ms;
  if (value_out) *value_out = *vsum * factor;
  if (g_trace.thr >= 0 && g_trace.ms(0, 3) - ms > g_trace.thr)
    fprintf(stderr, "[nfcuda host trace] plan %.2f ms, enqueue %.2f ms, wait %.2f ms, device %.2f ms\n", g_trace.ms(0, 1),
            g_trace.ms(1, 2), g_trace.ms(2, 3), (double)ms);
  return NF_OK;
}

// Host rows -> the device buffer `d` of a value+gradient call.  Large batches on the layered path go in two halves: the first on
// the compute stream, the second on a copy stream while the first half is already being computed (the chunk loop of
// general.cu runs forward + backward per chunk and waits on in_ev before the second chunk).  Each half gets its own
// per-tensor scales, like any chunked run.  OFF by default (NFCUDA_H2D_OVERLAP=1 enables it): measured on B200, C3 at 2^20 with
// host draws, 18.22 M samples/s with the overlap against 18.63 M with one copy -- the first coupling needs the exact maximum of
// its whole input, so overlap needs two chunks, and two half-size passes cost more than half a 274 MB copy (2.5 ms) hides.
static int stage_host_input(Flow& f, void* d, const void* h, int64_t rows, size_t row_bytes) {
  static const bool overlap = getenv("NFCUDA_H2D_OVERLAP") && atoi(getenv("NFCUDA_H2D_OVERLAP")) != 0;
  f.in_ev_armed = false;
  if (!overlap || f.all_elementwise || rows < ((int64_t)1 << 18) || f.chunk_N < rows) {
    NF_CUDA(cudaMemcpyAsync(d, h, (size_t)rows * row_bytes, cudaMemcpyHostToDevice, f.stream));
    return NF_OK;
  }
  if (!f.copy_stream) NF_CUDA(cudaStreamCreateWithFlags(&f.copy_stream, cudaStreamNonBlocking));
  if (!f.in_ev) NF_CUDA(cudaEventCreateWithFlags(&f.in_ev, cudaEventDisableTiming));
  const int64_t split = round_up((rows + 1) / 2, 1024);
  f.chunk_N = split;
  NF_CUDA(cudaMemcpyAsync(d, h, (size_t)split * row_bytes, cudaMemcpyHostToDevice, f.stream));
  // (the buffer's previous readers finished before the previous call returned: every host-buffer call ends with a stream sync)
  NF_CUDA(cudaMemcpyAsync((char*)d + (size_t)split * row_bytes, (const char*)h + (size_t)split * row_bytes,
                          (size_t)(rows - split) * row_bytes, cudaMemcpyHostToDevice, f.copy_stream));
  NF_CUDA(cudaEventRecord(f.in_ev, f.copy_stream));
  f.in_ev_armed = true;
  return NF_OK;
}

static int value_and_grad_host(Flow& f, int op, const Target* tgt, const void* theta_host, int64_t N, const void* in_host,
                               uint64_t seed, double scale, double* value_out, void* grad_host_out) {
  NF_REQUIRE(theta_host, "theta is null");
  NF_REQUIRE(N > 0, "N must be positive");
  g_trace.mark(0);
  NF_CUDA(cudaSetDevice(f.device));
  const size_t es = f.esize();
  f.ws_reset();
  const size_t in_bytes = in_host ? (size_t)N * f.dim * es : 0;
  NF_TRY(general_plan_workspace(f, op, N, in_bytes));
  NF_CUDA(cudaMemcpyAsync(f.d_theta, theta_host, f.P * es, cudaMemcpyHostToDevice, f.stream));
  void* in_dev = nullptr;
  if (in_host) {
    in_dev = f.ws_alloc(in_bytes);
    if (!in_dev) return NF_ERR_OOM;
    NF_TRY(stage_host_input(f, in_dev, in_host, N, (size_t)f.dim * es));
  }
  NF_TRY(value_and_grad_dev(f, op, tgt, f.d_theta, N, in_dev, seed, scale, value_out, grad_host_out ? f.d_out : nullptr));
  if (grad_host_out) {
    NF_CUDA(cudaMemcpyAsync(f.h_pinned, f.d_out, f.P * es, cudaMemcpyDeviceToHost, f.stream));
    NF_CUDA(cudaStreamSynchronize(f.stream));
    memcpy(grad_host_out, f.h_pinned, f.P * es);
  }
  return NF_OK;
}

// forward / inverse / logpdf style calls with host buffers
static int transform_host(Flow& f, int op, const Target* tgt, const void* theta_host, int64_t N, const void* in_host,
                          uint64_t seed, void* y_host, void* ld_host, void* terms_host, int32_t* bins_host, size_t bins_count) {
  NF_REQUIRE(theta_host, "theta is null");
  NF_REQUIRE(N > 0, "N must be positive");
  NF_CUDA(cudaSetDevice(f.device));
  const size_t es = f.esize();
  f.ws_reset();
  const size_t mat = (size_t)N * f.dim * es, vec = (size_t)N * es;
  NF_TRY(general_plan_workspace(f, op, N, 2 * mat + 2 * vec + bins_count * sizeof(int32_t) + 4096));
  NF_CUDA(cudaMemcpyAsync(f.d_theta, theta_host, f.P * es, cudaMemcpyHostToDevice, f.stream));
  Job j;
  j.op = op; j.tgt = tgt; j.theta_dev = f.d_theta; j.N = N; j.seed = seed;
  if (in_host) {
    void* in_dev = f.ws_alloc(mat);
    if (!in_dev) return NF_ERR_OOM;
    NF_CUDA(cudaMemcpyAsync(in_dev, in_host, mat, cudaMemcpyHostToDevice, f.stream));
    j.in_dev = in_dev;
  }
  if (y_host) { j.y_out = f.ws_alloc(mat); if (!j.y_out) return NF_ERR_OOM; }
  if (ld_host) { j.ld_out = f.ws_alloc(vec); if (!j.ld_out) return NF_ERR_OOM; }
  if (terms_host) { j.terms_out = f.ws_alloc(vec); if (!j.terms_out) return NF_ERR_OOM; }
  if (bins_host) { j.bins_out = (int32_t*)f.ws_alloc(bins_count * sizeof(int32_t)); if (!j.bins_out) return NF_ERR_OOM; }
  NF_TRY(run_job(f, j));
  if (y_host) NF_CUDA(cudaMemcpyAsync(y_host, j.y_out, mat, cudaMemcpyDeviceToHost, f.stream));
  if (ld_host) NF_CUDA(cudaMemcpyAsync(ld_host, j.ld_out, vec, cudaMemcpyDeviceToHost, f.stream));
  if (terms_host) NF_CUDA(cudaMemcpyAsync(terms_host, j.terms_out, vec, cudaMemcpyDeviceToHost, f.stream));
  if (bins_host) NF_CUDA(cudaMemcpyAsync(bins_host, j.bins_out, bins_count * sizeof(int32_t), cudaMemcpyDeviceToHost, f.stream));
  NF_CUDA(cudaStreamSynchronize(f.stream));
  return NF_OK;
}

}  // namespace nf

using namespace nf;

#define NF_FLOW(h) (*reinterpret_cast<nf::Flow*>(h))
#define NF_FLOW_REF(h) (*reinterpret_cast<nf::Flow*>(h))
#define NF_CHECK_HANDLE(h)                                  \
  do {                                                      \
    if (!(h)) { nf::set_error("null handle"); return NF_ERR_INVALID; } \
  } while (0)

extern "C" {

int nf_version(void) { return NFCUDA_VERSION; }
const char* nf_last_error(void) { return g_last_error.c_str(); }

int nf_device_count(int* count) {
  NF_REQUIRE(count, "null argument");
  NF_CUDA(cudaGetDeviceCount(count));
  return NF_OK;
}

int nf_init(int device) {
  int n = 0;
  NF_CUDA(cudaGetDeviceCount(&n));
  NF_REQUIRE(device >= 0 && device < n, "device %d out of range (found %d CUDA devices)", device, n);
  NF_CUDA(cudaSetDevice(device));
  return check_device();
}

int nf_synchronize(void) {
  NF_CUDA(cudaDeviceSynchronize());
  return NF_OK;
}

int nf_flow_create(nf_flow_t* out, const nf_layer_desc* layers, int n_layers, int dim, int dtype) {
  return flow_create(out, layers, n_layers, dim, dtype);
}
void nf_flow_destroy(nf_flow_t flow) { flow_destroy(reinterpret_cast<Flow*>(flow)); }
int64_t nf_flow_num_params(nf_flow_t flow) { return flow ? NF_FLOW(flow).P : -1; }
int nf_flow_dim(nf_flow_t flow) { return flow ? NF_FLOW(flow).dim : -1; }
int64_t nf_flow_param_offset(nf_flow_t flow, int layer) {
  if (!flow || layer < 0 || layer >= (int)NF_FLOW(flow).layers.size()) return -1;
  return NF_FLOW(flow).layers[layer].theta_off;
}

int nf_flow_set_base(nf_flow_t flow, const double* mu, const double* sigma) {
  NF_CHECK_HANDLE(flow);
  Flow& f = NF_FLOW(flow);
  NF_CUDA(cudaSetDevice(f.device));
  for (int k = 0; k < f.dim; ++k) {
    f.base_mu[k] = mu ? mu[k] : 0.0;
    f.base_sigma[k] = sigma ? sigma[k] : 1.0;
    NF_REQUIRE(f.base_sigma[k] > 0, "base sigma must be positive");
  }
  f.base_dense = false;
  return upload_base(f);
}

int nf_flow_set_base_chol(nf_flow_t flow, const double* mu, const double* L) {
  NF_CHECK_HANDLE(flow);
  Flow& f = NF_FLOW(flow);
  NF_REQUIRE(L, "nf_flow_set_base_chol: L is null");
  NF_CUDA(cudaSetDevice(f.device));
  const int d = f.dim;
  double c0 = -0.5 * d * NF_LOG2PI;
  for (int i = 0; i < d; ++i) {
    NF_REQUIRE(L[i * d + i] > 0, "nf_flow_set_base_chol: the Cholesky factor needs a positive diagonal (L[%d][%d] = %g)", i, i, L[i * d + i]);
    c0 -= std::log(L[i * d + i]);
  }
  const size_t n = (size_t)d * d + d;
  if (!f.d_base_L) NF_CUDA(cudaMalloc(&f.d_base_L, n * sizeof(double)));
  std::vector<double> h(n, 0.0);
  for (int i = 0; i < d; ++i)
    for (int k = 0; k <= i; ++k) h[(size_t)i * d + k] = L[i * d + k];     // strictly upper part is ignored
  for (int i = 0; i < d; ++i) { h[(size_t)d * d + i] = mu ? mu[i] : 0.0; f.base_mu[i] = h[(size_t)d * d + i]; }
  if (f.dtype == NF_F32) {
    std::vector<float> hf(h.begin(), h.end());
    NF_CUDA(cudaMemcpy(f.d_base_L, hf.data(), n * sizeof(float), cudaMemcpyHostToDevice));
  } else {
    NF_CUDA(cudaMemcpy(f.d_base_L, h.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  }
  f.base_dense = true;
  f.base_is_standard = false;
  f.base_c0 = c0;
  f.plan_valid = false;
  return NF_OK;
}

int nf_flow_set_mma_mode(nf_flow_t flow, int mode) {
  NF_CHECK_HANDLE(flow);
  Flow& f = NF_FLOW(flow);
  NF_REQUIRE(mode == NF_MMA_SIMT || mode == NF_MMA_F16X3 || mode == NF_MMA_F16X1, "unknown mma mode %d", mode);
  NF_REQUIRE(f.dtype == NF_F32 || mode == NF_MMA_SIMT, "Float64 flows only support NF_MMA_SIMT");
  f.mma_mode = mode;
  return NF_OK;
}

int nf_flow_set_workspace_limit(nf_flow_t flow, size_t bytes) {
  NF_CHECK_HANDLE(flow);
  NF_REQUIRE(bytes >= ((size_t)64 << 20), "workspace limit must be at least 64 MiB");
  NF_FLOW(flow).ws_limit = bytes;
  return NF_OK;
}

int nf_target_create(nf_target_t* out, int kind, int dim, const double* params, int n_params) {
  NF_REQUIRE(out, "null argument");
  NF_REQUIRE(dim > 0, "dim must be positive");
  std::unique_ptr<Target> t(new Target());
  t->kind = kind; t->dim = dim;
  for (int i = 0; i < n_params; ++i) t->p.push_back(params[i]);
  switch (kind) {
    case NF_TARGET_BANANA:
      NF_REQUIRE(n_params == 2 && dim >= 2 && params[1] > 0, "Banana(dim>=2; b, var>0)");
      t->c0 = -(std::log(params[1]) / dim + NF_LOG2PI) * dim / 2;
      break;
    case NF_TARGET_FUNNEL:
      NF_REQUIRE(n_params == 2 && dim >= 2 && params[1] > 0, "Funnel(dim>=2; mu, sigma>0)");
      t->c0 = -0.5 * NF_LOG2PI - std::log(params[1]) - (dim - 1) / 2.0 * NF_LOG2PI;
      break;
    case NF_TARGET_WARPED_GAUSS:
      NF_REQUIRE(n_params == 2 && dim == 2 && params[0] > 0 && params[1] > 0, "WarpedGauss(dim=2; sigma1>0, sigma2>0)");
      t->c0 = -NF_LOG2PI - std::log(params[0]) - std::log(params[1]);
      break;
    case NF_TARGET_CROSS:
      NF_REQUIRE(n_params == 2 && dim % 2 == 0 && params[1] > 0, "Cross(dim even; mu, sigma>0)");
      t->c0 = (dim / 2) * (std::log(0.25) - NF_LOG2PI - std::log(params[1]));
      break;
    case NF_TARGET_DIAG_NORMAL: {
      NF_REQUIRE(n_params == 2 * dim, "DiagNormal needs mu[dim], sigma[dim]");
      double c0 = -0.5 * dim * NF_LOG2PI;
      std::vector<float> hf(2 * dim);
      for (int k = 0; k < dim; ++k) {
        NF_REQUIRE(params[dim + k] > 0, "sigma must be positive");
        c0 -= std::log(params[dim + k]);
      }
      for (int k = 0; k < 2 * dim; ++k) hf[k] = (float)params[k];
      t->c0 = c0;
      NF_CUDA(cudaMalloc(&t->d_vec_f32, 2 * dim * sizeof(float)));
      NF_CUDA(cudaMalloc(&t->d_vec_f64, 2 * dim * sizeof(double)));
      NF_CUDA(cudaMemcpy(t->d_vec_f32, hf.data(), 2 * dim * sizeof(float), cudaMemcpyHostToDevice));
      NF_CUDA(cudaMemcpy(t->d_vec_f64, params, 2 * dim * sizeof(double), cudaMemcpyHostToDevice));
      break;
    }
    case NF_TARGET_LOGREG: {
      NF_REQUIRE(n_params >= 2 && params[0] > 0 && params[1] >= 1 && dim <= NF_LOGREG_MAX_DIM,
                 "LogReg(dim <= %d; sigma0 > 0, n >= 1, X[n*dim], y[n])", NF_LOGREG_MAX_DIM);
      const int n = (int)params[1];
      NF_REQUIRE(n_params == 2 + n * dim + n, "LogReg needs 2 + n*dim + n parameters, got %d", n_params);
      t->n_data = n;
      t->c0 = -0.5 * dim * (NF_LOG2PI + 2.0 * std::log(params[0]));
      const int cnt = n * dim + n;
      std::vector<float> hf(cnt);
      for (int k = 0; k < cnt; ++k) hf[k] = (float)params[2 + k];
      NF_CUDA(cudaMalloc(&t->d_vec_f32, cnt * sizeof(float)));
      NF_CUDA(cudaMalloc(&t->d_vec_f64, cnt * sizeof(double)));
      NF_CUDA(cudaMemcpy(t->d_vec_f32, hf.data(), cnt * sizeof(float), cudaMemcpyHostToDevice));
      NF_CUDA(cudaMemcpy(t->d_vec_f64, params + 2, cnt * sizeof(double), cudaMemcpyHostToDevice));
      t->p.resize(1);      // p[0] = sigma0 (the data live on the device)
      break;
    }
    default:
      set_error("unknown target kind %d", kind);
      return NF_ERR_INVALID;
  }
  *out = reinterpret_cast<nf_target_t>(t.release());
  return NF_OK;
}

int nf_target_create_joint(nf_target_t* out, nf_target_t inner) {
  NF_REQUIRE(out && inner, "null argument");
  const Target* in = reinterpret_cast<const Target*>(inner);
  NF_REQUIRE(!in->joint, "inner target is already a joint target");
  std::unique_ptr<Target> t(new Target());
  NF_TRY(clone_target(*in, *t));
  t->joint = true;
  t->dim = 2 * in->dim;
  *out = reinterpret_cast<nf_target_t>(t.release());
  return NF_OK;
}

int nf_target_logp(nf_target_t target, int dtype, const void* x_host, int64_t N, void* logp_host_out, void* score_host_out) {
  const Target* t = reinterpret_cast<const Target*>(target);
  NF_REQUIRE(t && x_host && logp_host_out && N > 0, "nf_target_logp: null argument / N must be positive");
  NF_REQUIRE(dtype == NF_F32 || dtype == NF_F64, "dtype must be NF_F32 or NF_F64");
  NF_REQUIRE(!t->joint, "nf_target_logp: joint [x, rho] targets are evaluated through their inner target");
  NF_TRY(check_device());
  const size_t es = dtype == NF_F64 ? 8 : 4;
  const int d = t->dim;
  void *dx = nullptr, *dl = nullptr, *dg = nullptr;
  auto body = [&]() -> int {
    NF_CUDA(cudaMalloc(&dx, (size_t)N * d * es));
    NF_CUDA(cudaMalloc(&dl, (size_t)N * es));
    NF_CUDA(cudaMalloc(&dg, (size_t)N * d * es));
    NF_CUDA(cudaMemcpy(dx, x_host, (size_t)N * d * es, cudaMemcpyHostToDevice));
    NF_CUDA(cudaMemset(dg, 0, (size_t)N * d * es));
    if (dtype == NF_F32) target_logp_kernel<float><<<(unsigned)ceil_div(N, 128), 128>>>(t->params<float>(), (const float*)dx, N, (float*)dl, (float*)dg);
    else target_logp_kernel<double><<<(unsigned)ceil_div(N, 128), 128>>>(t->params<double>(), (const double*)dx, N, (double*)dl, (double*)dg);
    NF_LAUNCH_CHECK();
    NF_CUDA(cudaMemcpy(logp_host_out, dl, (size_t)N * es, cudaMemcpyDeviceToHost));
    if (score_host_out) NF_CUDA(cudaMemcpy(score_host_out, dg, (size_t)N * d * es, cudaMemcpyDeviceToHost));
    return NF_OK;
  };
  const int s = body();
  cudaFree(dx); cudaFree(dl); cudaFree(dg);
  return s;
}

void nf_target_destroy(nf_target_t target) {
  Target* t = reinterpret_cast<Target*>(target);
  if (!t) return;
  cudaFree(t->d_vec_f32); cudaFree(t->d_vec_f64);
  delete t;
}

static int check_target(const Flow& f, const Target* t) {
  NF_REQUIRE(t, "null target");
  NF_REQUIRE(t->dim == f.dim, "target dim %d != flow dim %d", t->dim, f.dim);
  if (t->joint && !(f.all_elementwise && f.dim == 2 * (t->dim / 2) &&
                    (((f.dim & (f.dim - 1)) == 0 && f.dim <= 64) || hmc_warp_qualifies(f, t)))) {
    set_error("joint [x, rho] targets are implemented for elementwise / Hamiltonian flows with dim a power of two <= 64, and for "
              "flows of Shift / Scale / momentum-affine / LeapFrog layers with dim = 2h, h <= 128 (LogReg, Funnel, Banana, DiagNormal targets)");
    return NF_ERR_UNSUPPORTED;
  }
  return NF_OK;
}

int nf_elbo_value_and_grad(nf_flow_t flow, nf_target_t target, const void* theta_host, int64_t N, const void* z0_host,
                           uint64_t seed, double scale, double* value_out, void* grad_host_out) {
  NF_CHECK_HANDLE(flow);
  Flow& f = NF_FLOW(flow);
  const Target* t = reinterpret_cast<Target*>(target);
  NF_TRY(check_target(f, t));
  return value_and_grad_host(f, OP_ELBO, t, theta_host, N, z0_host, seed, scale, value_out, grad_host_out);
}

int nf_elbo_value_and_grad_dev(nf_flow_t flow, nf_target_t target, const void* theta_dev, int64_t N, const void* z0_dev,
                               uint64_t seed, double scale, double* value_out, void* grad_dev_out) {
  NF_CHECK_HANDLE(flow);
  Flow& f = NF_FLOW(flow);
  const Target* t = reinterpret_cast<Target*>(target);
  NF_TRY(check_target(f, t));
  NF_REQUIRE(theta_dev, "theta is null");
  g_trace.mark(0);
  NF_CUDA(cudaSetDevice(f.device));
  f.ws_reset();
  NF_TRY(general_plan_workspace(f, OP_ELBO, N, 0));
  return value_and_grad_dev(f, OP_ELBO, t, theta_dev, N, z0_dev, seed, scale, value_out, grad_dev_out);
}

int nf_elbo_sums_dev(nf_flow_t flow, nf_target_t target, const void* theta_dev, int64_t N, const void* z0_dev,
                     uint64_t seed, void* sums_dev_out) {
  NF_CHECK_HANDLE(flow);
  Flow& f = NF_FLOW(flow);
  const Target* t = reinterpret_cast<Target*>(target);
  NF_TRY(check_target(f, t));
  NF_REQUIRE(theta_dev && sums_dev_out, "null argument");
  NF_CUDA(cudaSetDevice(f.device));
  f.ws_reset();
  NF_TRY(general_plan_workspace(f, OP_ELBO, N, 0));
  Job j;
  j.op = OP_ELBO; j.tgt = t; j.theta_dev = theta_dev; j.in_dev = z0_dev; j.N = N; j.seed = seed; j.want_grad = true;
  NF_CUDA(cudaEventRecord(f.ev0, f.stream));
  NF_TRY(run_job(f, j));
  NF_TRY(scale_outputs(f, 1.0, sums_dev_out, f.P + 1));
  NF_CUDA(cudaEventRecord(f.ev1, f.stream));
  NF_CUDA(cudaStreamSynchronize(f.stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, f.ev0, f.ev1);
  f.last_ms = ms;
  return NF_OK;
}

int nf_elbo_terms(nf_flow_t flow, nf_target_t target, const void* theta_host, int64_t N, const void* z0_host,
                  void* elbos_host_out) {
  NF_CHECK_HANDLE(flow);
  Flow& f = NF_FLOW(flow);
  const Target* t = reinterpret_cast<Target*>(target);
  NF_TRY(check_target(f, t));
  NF_REQUIRE(z0_host && elbos_host_out, "null argument");
  return transform_host(f, OP_ELBO, t, theta_host, N, z0_host, 0, nullptr, nullptr, elbos_host_out, nullptr, 0);
}

int nf_loglik_value_and_grad(nf_flow_t flow, const void* theta_host, int64_t N, const void* xs_host, double scale,
                             double* value_out, void* grad_host_out) {
  NF_CHECK_HANDLE(flow);
  NF_REQUIRE(xs_host, "xs is null");
  return value_and_grad_host(NF_FLOW(flow), OP_LOGLIK, nullptr, theta_host, N, xs_host, 0, scale, value_out, grad_host_out);
}

int nf_loglik_value_and_grad_dev(nf_flow_t flow, const void* theta_dev, int64_t N, const void* xs_dev, double scale,
                                 double* value_out, void* grad_dev_out) {
  NF_CHECK_HANDLE(flow);
  Flow& f = NF_FLOW(flow);
  NF_REQUIRE(theta_dev && xs_dev, "null argument");
  NF_CUDA(cudaSetDevice(f.device));
  f.ws_reset();
  NF_TRY(general_plan_workspace(f, OP_LOGLIK, N, 0));
  return value_and_grad_dev(f, OP_LOGLIK, nullptr, theta_dev, N, xs_dev, 0, scale, value_out, grad_dev_out);
}

static int train_graph(Flow& f, const Target* t, int64_t N, uint64_t seed, int n_iters, int t0, double eta, double beta1, double beta2,
                       double eps, char* dm, char* dv) {
  const int64_t P = f.P;
  if (!f.d_iter) {
    NF_CUDA(cudaMalloc((void**)&f.d_iter, sizeof(int64_t) + sizeof(unsigned int) * 2));
  }
  NF_CUDA(cudaMemsetAsync(f.d_iter, 0, sizeof(int64_t) + sizeof(unsigned int) * 2, f.stream));
  unsigned int* d_done = reinterpret_cast<unsigned int*>(f.d_iter + 1);
  // one untimed warm-up enqueue outside capture: sizes the workspace, builds tensor maps, sets function attributes
  f.ws_reset();
  NF_TRY(general_plan_workspace(f, OP_ELBO, N, 0));
  Job j;
  j.op = OP_ELBO; j.tgt = t; j.theta_dev = f.d_theta; j.in_dev = nullptr; j.N = N; j.seed = seed; j.seed_iter_dev = f.d_iter; j.want_grad = true;
  auto enqueue_iteration = [&]() -> int {
    f.ws_reset();
    NF_TRY(general_plan_workspace(f, OP_ELBO, N, 0));
    NF_TRY(run_job(f, j));
    const int threads = 256;
    const int blocks = (int)ceil_div(P, threads);
    if (f.dtype == NF_F32)
      adam_step_graph_kernel<float><<<blocks, threads, 0, f.stream>>>(f.d_gsum, P, 1.0 / (double)N, beta1, beta2, t0, (float)eta, (float)eps,
                                                                     (float*)f.d_theta, (float*)dm, (float*)dv, f.d_stats, f.d_iter, d_done);
    else
      adam_step_graph_kernel<double><<<blocks, threads, 0, f.stream>>>(f.d_gsum, P, 1.0 / (double)N, beta1, beta2, t0, eta, eps,
                                                                      (double*)f.d_theta, (double*)dm, (double*)dv, f.d_stats, f.d_iter, d_done);
    NF_LAUNCH_CHECK();
    return NF_OK;
  };
  NF_TRY(enqueue_iteration());                       // iteration 0, eagerly (also the warm-up)
  const void* ws_base = f.ws.base;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  NF_CUDA(cudaStreamBeginCapture(f.stream, cudaStreamCaptureModeThreadLocal));
  const int rc = enqueue_iteration();
  const cudaError_t ce = cudaStreamEndCapture(f.stream, &graph);
  if (rc != NF_OK || ce != cudaSuccess || !graph || f.ws.base != ws_base) {
    if (graph) cudaGraphDestroy(graph);
    set_error("CUDA graph capture of the training iteration failed");
    // iteration 0 has already been applied eagerly: finish the remaining iterations eagerly too
    for (int it = 1; it < n_iters; ++it) NF_TRY(enqueue_iteration());
    return NF_OK;
  }
  if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
    cudaGraphDestroy(graph);
    for (int it = 1; it < n_iters; ++it) NF_TRY(enqueue_iteration());
    return NF_OK;
  }
  for (int it = 1; it < n_iters; ++it) {
    if (cudaGraphLaunch(exec, f.stream) != cudaSuccess) { cudaGraphExecDestroy(exec); cudaGraphDestroy(graph); set_error("cudaGraphLaunch failed"); return NF_ERR_CUDA; }
  }
  NF_CUDA(cudaStreamSynchronize(f.stream));
  cudaGraphExecDestroy(exec);
  cudaGraphDestroy(graph);
  return NF_OK;
}

int nf_train_elbo_adam(nf_flow_t flow, nf_target_t target, void* theta_host_inout, int64_t N, uint64_t seed, int n_iters,
                       int t0, double eta, double beta1, double beta2, double eps, void* m_host_inout, void* v_host_inout,
                       double* stats_out) {
  NF_CHECK_HANDLE(flow);
  Flow& f = NF_FLOW(flow);
  const Target* t = reinterpret_cast<Target*>(target);
  NF_TRY(check_target(f, t));
  NF_REQUIRE(theta_host_inout && stats_out && n_iters > 0 && N > 0 && t0 >= 0, "bad argument");
  NF_CUDA(cudaSetDevice(f.device));
  const size_t es = f.esize();
  const int64_t P = f.P;
  if (!f.d_adam) {
    NF_CUDA(cudaMalloc(&f.d_adam, 2 * P * es));
  }
  if (f.stats_cap < n_iters) {
    if (f.d_stats) cudaFree(f.d_stats);
    NF_CUDA(cudaMalloc((void**)&f.d_stats, (size_t)n_iters * 2 * sizeof(double)));
    f.stats_cap = n_iters;
  }
  char* dm = (char*)f.d_adam;
  char* dv = dm + P * es;
  NF_CUDA(cudaMemcpyAsync(f.d_theta, theta_host_inout, P * es, cudaMemcpyHostToDevice, f.stream));
  if (m_host_inout && v_host_inout && t0 > 0) {
    NF_CUDA(cudaMemcpyAsync(dm, m_host_inout, P * es, cudaMemcpyHostToDevice, f.stream));
    NF_CUDA(cudaMemcpyAsync(dv, v_host_inout, P * es, cudaMemcpyHostToDevice, f.stream));
  } else {
    NF_CUDA(cudaMemsetAsync(dm, 0, 2 * P * es, f.stream));
  }
  NF_CUDA(cudaMemsetAsync(f.d_stats, 0, (size_t)n_iters * 2 * sizeof(double), f.stream));
  NF_CUDA(cudaEventRecord(f.ev0, f.stream));
  // small batches of elementwise flows: every iteration inside one persistent single-CTA launch (NFCUDA_TRAIN_PERSISTENT=0 disables)
  const bool persistent_ok = !(getenv("NFCUDA_TRAIN_PERSISTENT") && atoi(getenv("NFCUDA_TRAIN_PERSISTENT")) == 0);
  bool persistent = persistent_ok && f.all_elementwise && f.dim <= 16 && N <= 2048;
  if (persistent) {
    // NF_ERR_UNSUPPORTED = "does not qualify" (e.g. the layer tables of a very deep flow exceed the shared memory of one
    // CTA): fall through to the multi-launch loop below instead of failing the call
    const int s = f.dtype == NF_F32 ? ew_train<float>(f, t, N, seed, n_iters, t0, eta, beta1, beta2, eps, dm, dv)
                                    : ew_train<double>(f, t, N, seed, n_iters, t0, eta, beta1, beta2, eps, dm, dv);
    if (s == NF_ERR_UNSUPPORTED) persistent = false;
    else NF_TRY(s);
  }
  // coupling flows with small batches are launch bound (~100 launches per iteration): capture ONE iteration (value + gradient +
  // Adam, the iteration index and the Philox seed offset read from a device counter) in a CUDA graph and replay it
  // (NFCUDA_TRAIN_GRAPH=0 disables; a capture problem makes train_graph finish with eager launches)
  bool graphed = false;
  if (!persistent && !f.all_elementwise && !f.prof.on && n_iters >= 4 && N <= ((int64_t)1 << 16) &&
      !(getenv("NFCUDA_TRAIN_GRAPH") && atoi(getenv("NFCUDA_TRAIN_GRAPH")) == 0)) {
    NF_TRY(train_graph(f, t, N, seed, n_iters, t0, eta, beta1, beta2, eps, dm, dv));   // falls back to eager launches inside
    graphed = true;
  }
  for (int it = 0; it < n_iters && !persistent && !graphed; ++it) {
    f.ws_reset();
    NF_TRY(general_plan_workspace(f, OP_ELBO, N, 0));
    Job j;
    j.op = OP_ELBO; j.tgt = t; j.theta_dev = f.d_theta; j.in_dev = nullptr; j.N = N; j.seed = seed + (uint64_t)it; j.want_grad = true;
    NF_TRY(run_job(f, j));
    const int step = t0 + it + 1;
    const double omb1t = 1.0 - std::pow(beta1, step), omb2t = 1.0 - std::pow(beta2, step);
    const int threads = 256;
    const int blocks = (int)ceil_div(P, threads);
    if (f.dtype == NF_F32)
      adam_step_kernel<float><<<blocks, threads, 0, f.stream>>>(f.d_gsum, P, 1.0 / (double)N, (float)beta1, (float)beta2, (float)omb1t,
                                                               (float)omb2t, (float)eta, (float)eps, (float*)f.d_theta, (float*)dm,
                                                               (float*)dv, f.d_stats + 2 * it);
    else
      adam_step_kernel<double><<<blocks, threads, 0, f.stream>>>(f.d_gsum, P, 1.0 / (double)N, beta1, beta2, omb1t, omb2t, eta, eps,
                                                                (double*)f.d_theta, (double*)dm, (double*)dv, f.d_stats + 2 * it);
    NF_LAUNCH_CHECK();
  }
  NF_CUDA(cudaEventRecord(f.ev1, f.stream));
  NF_CUDA(cudaMemcpyAsync(theta_host_inout, f.d_theta, P * es, cudaMemcpyDeviceToHost, f.stream));
  if (m_host_inout && v_host_inout) {
    NF_CUDA(cudaMemcpyAsync(m_host_inout, dm, P * es, cudaMemcpyDeviceToHost, f.stream));
    NF_CUDA(cudaMemcpyAsync(v_host_inout, dv, P * es, cudaMemcpyDeviceToHost, f.stream));
  }
  NF_CUDA(cudaMemcpyAsync(stats_out, f.d_stats, (size_t)n_iters * 2 * sizeof(double), cudaMemcpyDeviceToHost, f.stream));
  NF_CUDA(cudaStreamSynchronize(f.stream));
  for (int it = 0; it < n_iters; ++it) stats_out[2 * it + 1] = std::sqrt(stats_out[2 * it + 1]);
  float ms = 0;
  cudaEventElapsedTime(&ms, f.ev0, f.ev1);
  f.last_ms = ms;
  return NF_OK;
}

int nf_forward(nf_flow_t flow, const void* theta_host, int64_t N, const void* x_host, void* y_host_out, void* logdet_host_out) {
  NF_CHECK_HANDLE(flow);
  NF_REQUIRE(x_host, "x is null");
  return transform_host(NF_FLOW(flow), OP_FORWARD, nullptr, theta_host, N, x_host, 0, y_host_out, logdet_host_out, nullptr, nullptr, 0);
}

int nf_inverse(nf_flow_t flow, const void* theta_host, int64_t N, const void* y_host, void* x_host_out, void* logdet_host_out) {
  NF_CHECK_HANDLE(flow);
  NF_REQUIRE(y_host, "y is null");
  return transform_host(NF_FLOW(flow), OP_INVERSE, nullptr, theta_host, N, y_host, 0, x_host_out, logdet_host_out, nullptr, nullptr, 0);
}

int nf_logpdf(nf_flow_t flow, const void* theta_host, int64_t N, const void* y_host, void* logpdf_host_out) {
  NF_CHECK_HANDLE(flow);
  NF_REQUIRE(y_host && logpdf_host_out, "null argument");
  return transform_host(NF_FLOW(flow), OP_LOGLIK, nullptr, theta_host, N, y_host, 0, nullptr, nullptr, logpdf_host_out, nullptr, 0);
}

int nf_sample(nf_flow_t flow, const void* theta_host, int64_t N, uint64_t seed, void* y_host_out) {
  NF_CHECK_HANDLE(flow);
  NF_REQUIRE(y_host_out, "null argument");
  return transform_host(NF_FLOW(flow), OP_FORWARD, nullptr, theta_host, N, nullptr, seed, y_host_out, nullptr, nullptr, nullptr, 0);
}

int nf_base_sample(nf_flow_t flow, int64_t N, uint64_t seed, void* z_host_out) {
  NF_CHECK_HANDLE(flow);
  Flow& f = NF_FLOW(flow);
  NF_REQUIRE(z_host_out && N > 0, "bad argument");
  NF_CUDA(cudaSetDevice(f.device));
  f.ws_reset();
  const size_t mat = (size_t)N * f.dim * f.esize();
  NF_TRY(f.ws_reserve(mat + 4096));
  void* z = f.ws_alloc(mat);
  if (!z) return NF_ERR_OOM;
  NF_TRY(base_sample_dev(f, N, seed, z));
  NF_CUDA(cudaMemcpyAsync(z_host_out, z, mat, cudaMemcpyDeviceToHost, f.stream));
  NF_CUDA(cudaStreamSynchronize(f.stream));
  return NF_OK;
}

int nf_forward_stash(nf_flow_t flow, const void* theta_host, int64_t N, const void* x_host, void* y_host_out,
                     void* logdet_host_out) {
  NF_CHECK_HANDLE(flow);
  Flow& f = NF_FLOW(flow);
  NF_REQUIRE(x_host, "x is null");
  NF_REQUIRE(!f.all_elementwise, "two-phase API is implemented for coupling flows; elementwise flows use the fused ELBO kernel");
  int s = transform_host(f, OP_FORWARD_STASH, nullptr, theta_host, N, x_host, 0, y_host_out, logdet_host_out, nullptr, nullptr, 0);
  if (s == NF_OK) f.stash_N = N;
  return s;
}

int nf_backward(nf_flow_t flow, const void* gy_host, const void* gld_host_or_null, void* grad_host_out) {
  NF_CHECK_HANDLE(flow);
  Flow& f = NF_FLOW(flow);
  NF_REQUIRE(gy_host && grad_host_out, "null argument");
  NF_REQUIRE(f.stash_N > 0, "nf_backward needs a preceding nf_forward_stash");
  NF_CUDA(cudaSetDevice(f.device));
  const size_t es = f.esize();
  NF_TRY(general_backward_from_stash(f, gy_host, gld_host_or_null));
  NF_TRY(scale_outputs(f, 1.0, f.d_out, f.P));
  NF_CUDA(cudaMemcpyAsync(f.h_pinned, f.d_out, f.P * es, cudaMemcpyDeviceToHost, f.stream));
  NF_CUDA(cudaStreamSynchronize(f.stream));
  memcpy(grad_host_out, f.h_pinned, f.P * es);
  f.stash_N = 0;
  return NF_OK;
}

int nf_spline_bins(nf_flow_t flow, const void* theta_host, int64_t N, const void* x_host, int32_t* bins_out) {
  NF_CHECK_HANDLE(flow);
  Flow& f = NF_FLOW(flow);
  NF_REQUIRE(x_host && bins_out, "null argument");
  size_t count = 0;
  for (auto& L : f.layers) if (L.kind == NF_SPLINE_COUPLING) count += (size_t)N * L.idx1.size();
  NF_REQUIRE(count > 0, "flow has no spline couplings");
  return transform_host(f, OP_FORWARD, nullptr, theta_host, N, x_host, 0, nullptr, nullptr, nullptr, bins_out, count);
}

int nf_rqs_bin_search(int dtype, const void* knots_host, const void* v_host, int64_t M, int K, int32_t* bins_out) {
  NF_REQUIRE(knots_host && v_host && bins_out && M > 0 && K >= 1, "bad argument");
  NF_TRY(check_device());
  return rqs_bin_search_host(dtype, knots_host, v_host, M, K, bins_out);
}

int nf_tc_gemm_test(int64_t n, int K, int N, const float* X, const float* Wt, const float* b, int terms, float* Y) {
  NF_REQUIRE(X && Wt && b && Y && n > 0 && K > 0 && N > 0, "bad argument");
  NF_REQUIRE(terms == 1 || terms == 3, "terms must be 1 or 3");
  NF_TRY(check_device());
  return tc_gemm_selftest(n, K, N, X, Wt, b, terms, Y);
}

int nf_fused_schedule(int n_chunks, int chain, int n_hoist, int delay, unsigned char* items_out, int cap) {
  NF_REQUIRE(items_out && cap > 0 && n_chunks >= 1 && n_chunks <= 4 && (chain == 1 || chain == 2) && n_hoist >= std::min(chain, n_chunks) &&
                 delay >= 1,
             "nf_fused_schedule: n_chunks in 1..4, chain in {1, 2}, n_hoist >= min(chain, n_chunks), delay >= 1");
  return tc_fused_schedule(n_chunks, chain, std::min(n_hoist, n_chunks), delay, items_out, cap);
}

int64_t nf_launch_count(int reset) {
  const int64_t c = g_launch_count;
  if (reset) g_launch_count = 0;
  return c;
}

int nf_set_option(const char* name, int value) {
  if (name && !strcmp(name, "fused_coupling")) { g_opt_fused_coupling = value; return NF_OK; }
  if (name && !strcmp(name, "fused_variant")) { g_opt_fused_variant = value; return NF_OK; }
  if (name && !strcmp(name, "hmc_warp")) { g_opt_hmc_warp = value; return NF_OK; }
  if (name && !strcmp(name, "rqs_planes")) { g_opt_rqs_planes = value; return NF_OK; }
  set_error("nf_set_option: unknown option '%s'", name ? name : "(null)");
  return NF_ERR_INVALID;
}

double nf_last_device_ms(nf_flow_t flow) { return flow ? NF_FLOW(flow).last_ms : -1.0; }

int nf_profile_enable(nf_flow_t flow, int on) {
  NF_CHECK_HANDLE(flow);
  NF_FLOW(flow).prof.on = on != 0;
  return NF_OK;
}

int nf_profile_keys(nf_flow_t flow, char* buf, int buflen) {
  NF_CHECK_HANDLE(flow);
  NF_REQUIRE(buf && buflen > 0, "bad argument");
  const std::string k = NF_FLOW(flow).prof.keys();
  snprintf(buf, buflen, "%s", k.c_str());
  return NF_OK;
}

int nf_profile_collect(nf_flow_t flow, const char* key, int64_t* launches, double* total_ms) {
  NF_CHECK_HANDLE(flow);
  NF_REQUIRE(key && launches && total_ms, "null argument");
  Flow& f = NF_FLOW(flow);
  NF_CUDA(cudaSetDevice(f.device));
  NF_CUDA(cudaStreamSynchronize(f.stream));
  f.prof.collect(key, launches, total_ms);
  return NF_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// multi-GPU data parallelism: sample shards, replicated theta, ONE all-reduce of the P+1 accumulators
// (SURVEY 8e).  NCCL is bound at run time so that nothing else in the library depends on it.
// ---------------------------------------------------------------------------------------------
namespace nf {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
  if (g_nccl.lib) return NF_OK;
  // a process that already carries an NCCL (e.g. the one bundled with PyTorch) is joined to it; else the system library
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { set_error("NCCL is not available: %s", dlerror()); return NF_ERR_UNSUPPORTED; }
#define NF_NCCL_SYM(field, name)                                                     \
  do {                                                                               \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                       \
    if (!g_nccl.field) { set_error("NCCL symbol %s is missing", name); return NF_ERR_UNSUPPORTED; } \
  } while (0)
  NF_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
  NF_NCCL_SYM(CommInitAll, "ncclCommInitAll");
  NF_NCCL_SYM(CommInitRank, "ncclCommInitRank");
  NF_NCCL_SYM(CommDestroy, "ncclCommDestroy");
  NF_NCCL_SYM(AllReduce, "ncclAllReduce");
  NF_NCCL_SYM(GroupStart, "ncclGroupStart");
  NF_NCCL_SYM(GroupEnd, "ncclGroupEnd");
  NF_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NF_NCCL_SYM
  g_nccl.lib = h;
  return NF_OK;
}
#define NF_NCCL(expr)                                                                          \
  do {                                                                                         \
    ncclResult_t _r = (expr);                                                                  \
    if (_r != ncclSuccess) {                                                                   \
      nf::set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(_r)); \
      return NF_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

struct Comm {
  int n_ranks = 1;
  std::vector<int> devices;        // local devices, in rank order
  std::vector<int> ranks;          // global rank of each local device
  std::vector<ncclComm_t> comms;   // one per local device (empty for a single-rank job: nothing to reduce)
};

static void shard_range(int64_t N, int R, int r, int64_t* b, int64_t* e) {
  const int64_t q = N / R, rem = N % R;
  *b = r * q + std::min<int64_t>(r, rem);
  *e = *b + q + (r < rem ? 1 : 0);
}

// theta / inputs per local device are already resident (theta_dev[i], in_dev[i] or null); leaves the scaled gradient in
// each flow's d_out (and in grad_dev_out[i] when given) and the job-wide objective in *value_out.
static int multi_core(Comm& c, Flow* const* fl, const Target* const* tg, int op, const void* const* theta_dev, int64_t N_total,
                      const void* const* in_dev, uint64_t seed, double scale, double* value_out, void* const* grad_dev_out) {
  const int G = (int)c.devices.size();
  std::vector<int64_t> nloc(G);
  for (int i = 0; i < G; ++i) {
    Flow& f = *fl[i];
    int64_t b, e;
    shard_range(N_total, c.n_ranks, c.ranks[i], &b, &e);
    nloc[i] = e - b;
    NF_REQUIRE(nloc[i] > 0, "N_total = %lld leaves rank %d without samples", (long long)N_total, c.ranks[i]);
    NF_CUDA(cudaSetDevice(f.device));
    Job j;
    j.op = op; j.tgt = tg ? tg[i] : nullptr; j.theta_dev = theta_dev[i]; j.in_dev = in_dev ? in_dev[i] : nullptr; j.N = nloc[i]; j.seed = seed;
    j.want_grad = true;
    f.draw_row_offset = j.in_dev ? 0 : b;
    NF_CUDA(cudaEventRecord(f.ev0, f.stream));
    const int s = run_job(f, j);            // asynchronous: kernels queue on the device's stream, the host moves on to the next device
    f.draw_row_offset = 0;
    NF_TRY(s);
  }
  if (!c.comms.empty()) {
    NF_NCCL(g_nccl.GroupStart());
    for (int i = 0; i < G; ++i)
      NF_NCCL(g_nccl.AllReduce(fl[i]->d_gsum, fl[i]->d_gsum, (size_t)(fl[i]->P + 1), ncclDouble, ncclSum, c.comms[i], fl[i]->stream));
    NF_NCCL(g_nccl.GroupEnd());
  }
  const double factor = scale / (double)N_total;
  for (int i = 0; i < G; ++i) {
    Flow& f = *fl[i];
    NF_CUDA(cudaSetDevice(f.device));
    NF_TRY(scale_outputs(f, factor, grad_dev_out && grad_dev_out[i] ? grad_dev_out[i] : f.d_out, f.P));
    NF_CUDA(cudaEventRecord(f.ev1, f.stream));
  }
  Flow& f0 = *fl[0];
  NF_CUDA(cudaSetDevice(f0.device));
  double* vsum = reinterpret_cast<double*>((char*)f0.h_pinned + f0.h_pinned_bytes - 16);
  NF_CUDA(cudaMemcpyAsync(vsum, f0.d_gsum + f0.P, sizeof(double), cudaMemcpyDeviceToHost, f0.stream));
  for (int i = 0; i < G; ++i) {
    NF_CUDA(cudaSetDevice(fl[i]->device));
    NF_CUDA(cudaStreamSynchronize(fl[i]->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, fl[i]->ev0, fl[i]->ev1);
    fl[i]->last_ms = ms;
  }
  if (value_out) *value_out = *vsum * factor;
  return NF_OK;
}

static int multi_host(Comm& c, const nf_flow_t* flows, const nf_target_t* targets, int op, const void* theta_host, int64_t N_total,
                      const void* in_host, uint64_t seed, double scale, double* value_out, void* grad_host_out) {
  const int G = (int)c.devices.size();
  NF_REQUIRE(flows && theta_host && N_total > 0, "null argument / N_total must be positive");
  std::vector<Flow*> fl(G);
  std::vector<const Target*> tg(G, nullptr);
  std::vector<const void*> th(G), in(G, nullptr);
  int64_t first_row = 0, dummy = 0;
  shard_range(N_total, c.n_ranks, c.ranks[0], &first_row, &dummy);
  for (int i = 0; i < G; ++i) {
    NF_REQUIRE(flows[i], "null flow handle for local device %d", i);
    Flow& f = NF_FLOW_REF(flows[i]);
    fl[i] = &f;
    NF_REQUIRE(f.device == c.devices[i], "flows[%d] lives on device %d, the communicator expects device %d", i, f.device, c.devices[i]);
    NF_REQUIRE(f.P == fl[0]->P && f.dtype == fl[0]->dtype && f.dim == fl[0]->dim, "flows[%d] is not a replica of flows[0]", i);
    if (op == OP_ELBO) {
      NF_REQUIRE(targets && targets[i], "null target handle for local device %d", i);
      tg[i] = reinterpret_cast<const Target*>(targets[i]);
      NF_TRY(check_target(f, tg[i]));
    }
    int64_t b, e;
    shard_range(N_total, c.n_ranks, c.ranks[i], &b, &e);
    const size_t es = f.esize();
    NF_CUDA(cudaSetDevice(f.device));
    f.ws_reset();
    const size_t in_bytes = in_host ? (size_t)(e - b) * f.dim * es : 0;
    NF_TRY(general_plan_workspace(f, op, e - b, in_bytes));
    NF_CUDA(cudaMemcpyAsync(f.d_theta, theta_host, f.P * es, cudaMemcpyHostToDevice, f.stream));
    th[i] = f.d_theta;
    if (in_host) {
      void* d = f.ws_alloc(in_bytes);
      if (!d) return NF_ERR_OOM;
      NF_TRY(stage_host_input(f, d, (const char*)in_host + (size_t)(b - first_row) * f.dim * es, e - b, (size_t)f.dim * es));
      in[i] = d;
    }
  }
  NF_TRY(multi_core(c, fl.data(), op == OP_ELBO ? tg.data() : nullptr, op, th.data(), N_total, in_host ? in.data() : nullptr, seed, scale,
                    value_out, nullptr));
  if (grad_host_out) {
    Flow& f0 = *fl[0];
    const size_t es = f0.esize();
    NF_CUDA(cudaSetDevice(f0.device));
    NF_CUDA(cudaMemcpyAsync(f0.h_pinned, f0.d_out, f0.P * es, cudaMemcpyDeviceToHost, f0.stream));
    NF_CUDA(cudaStreamSynchronize(f0.stream));
    memcpy(grad_host_out, f0.h_pinned, f0.P * es);
  }
  return NF_OK;
}
}  // namespace nf

extern "C" {

void nf_shard_range(int64_t N_total, int n_ranks, int rank, int64_t* begin, int64_t* end) {
  int64_t b = 0, e = 0;
  if (n_ranks > 0 && rank >= 0 && rank < n_ranks) nf::shard_range(N_total, n_ranks, rank, &b, &e);
  if (begin) *begin = b;
  if (end) *end = e;
}

int nf_comm_init_all(nf_comm_t* out, int n_dev, const int* dev_ids) {
  NF_REQUIRE(out && n_dev >= 1, "nf_comm_init_all: need an output handle and n_dev >= 1");
  int count = 0;
  NF_CUDA(cudaGetDeviceCount(&count));
  std::unique_ptr<nf::Comm> c(new nf::Comm());
  c->n_ranks = n_dev;
  for (int i = 0; i < n_dev; ++i) {
    const int dev = dev_ids ? dev_ids[i] : i;
    NF_REQUIRE(dev >= 0 && dev < count, "nf_comm_init_all: device %d is not present (%d devices)", dev, count);
    c->devices.push_back(dev);
    c->ranks.push_back(i);
  }
  if (n_dev > 1) {
    NF_TRY(nf::nccl_load());
    c->comms.resize(n_dev);
    NF_NCCL(nf::g_nccl.CommInitAll(c->comms.data(), n_dev, c->devices.data()));
  }
  *out = reinterpret_cast<nf_comm_t>(c.release());
  return NF_OK;
}

int nf_comm_unique_id(void* id_out) {
  NF_REQUIRE(id_out, "null argument");
  static_assert(sizeof(ncclUniqueId) <= NF_UNIQUE_ID_BYTES, "NF_UNIQUE_ID_BYTES too small");
  NF_TRY(nf::nccl_load());
  ncclUniqueId id;
  NF_NCCL(nf::g_nccl.GetUniqueId(&id));
  memset(id_out, 0, NF_UNIQUE_ID_BYTES);
  memcpy(id_out, &id, sizeof(id));
  return NF_OK;
}

int nf_comm_init_rank(nf_comm_t* out, int n_ranks, int rank, const void* id, int device) {
  NF_REQUIRE(out && n_ranks >= 1 && rank >= 0 && rank < n_ranks, "nf_comm_init_rank: bad rank %d of %d", rank, n_ranks);
  int count = 0;
  NF_CUDA(cudaGetDeviceCount(&count));
  NF_REQUIRE(device >= 0 && device < count, "nf_comm_init_rank: device %d is not present (%d devices)", device, count);
  std::unique_ptr<nf::Comm> c(new nf::Comm());
  c->n_ranks = n_ranks;
  c->devices.push_back(device);
  c->ranks.push_back(rank);
  if (n_ranks > 1) {
    NF_REQUIRE(id, "nf_comm_init_rank: null unique id");
    NF_TRY(nf::nccl_load());
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    NF_CUDA(cudaSetDevice(device));
    c->comms.resize(1);
    NF_NCCL(nf::g_nccl.CommInitRank(&c->comms[0], n_ranks, uid, rank));
  }
  *out = reinterpret_cast<nf_comm_t>(c.release());
  return NF_OK;
}

int nf_comm_size(nf_comm_t comm) { return comm ? reinterpret_cast<nf::Comm*>(comm)->n_ranks : -1; }
int nf_comm_local_size(nf_comm_t comm) { return comm ? (int)reinterpret_cast<nf::Comm*>(comm)->devices.size() : -1; }
int nf_comm_local_device(nf_comm_t comm, int i) {
  if (!comm) return -1;
  auto* c = reinterpret_cast<nf::Comm*>(comm);
  return (i >= 0 && i < (int)c->devices.size()) ? c->devices[i] : -1;
}
int nf_comm_local_rank(nf_comm_t comm, int i) {
  if (!comm) return -1;
  auto* c = reinterpret_cast<nf::Comm*>(comm);
  return (i >= 0 && i < (int)c->ranks.size()) ? c->ranks[i] : -1;
}

void nf_comm_destroy(nf_comm_t comm) {
  if (!comm) return;
  auto* c = reinterpret_cast<nf::Comm*>(comm);
  for (size_t i = 0; i < c->comms.size(); ++i) {
    cudaSetDevice(c->devices[i]);
    if (nf::g_nccl.CommDestroy) nf::g_nccl.CommDestroy(c->comms[i]);
  }
  delete c;
}

int nf_elbo_value_and_grad_multi(nf_comm_t comm, const nf_flow_t* flows, const nf_target_t* targets, const void* theta_host,
                                 int64_t N_total, const void* z0_host, uint64_t seed, double scale, double* value_out,
                                 void* grad_host_out) {
  NF_REQUIRE(comm, "null communicator");
  return nf::multi_host(*reinterpret_cast<nf::Comm*>(comm), flows, targets, OP_ELBO, theta_host, N_total, z0_host, seed, scale, value_out,
                        grad_host_out);
}

int nf_loglik_value_and_grad_multi(nf_comm_t comm, const nf_flow_t* flows, const void* theta_host, int64_t N_total,
                                   const void* xs_host, double scale, double* value_out, void* grad_host_out) {
  NF_REQUIRE(comm, "null communicator");
  NF_REQUIRE(xs_host, "loglikelihood needs data");
  return nf::multi_host(*reinterpret_cast<nf::Comm*>(comm), flows, nullptr, OP_LOGLIK, theta_host, N_total, xs_host, 0, scale, value_out,
                        grad_host_out);
}

int nf_elbo_value_and_grad_multi_dev(nf_comm_t comm, const nf_flow_t* flows, const nf_target_t* targets, const void* const* theta_dev,
                                     int64_t N_total, const void* const* z0_dev, uint64_t seed, double scale, double* value_out,
                                     void* const* grad_dev_out) {
  NF_REQUIRE(comm && flows && targets && theta_dev && N_total > 0, "null argument / N_total must be positive");
  nf::Comm& c = *reinterpret_cast<nf::Comm*>(comm);
  const int G = (int)c.devices.size();
  std::vector<nf::Flow*> fl(G);
  std::vector<const nf::Target*> tg(G);
  for (int i = 0; i < G; ++i) {
    NF_REQUIRE(flows[i] && targets[i] && theta_dev[i], "null handle / theta for local device %d", i);
    nf::Flow& f = NF_FLOW_REF(flows[i]);
    fl[i] = &f;
    tg[i] = reinterpret_cast<const nf::Target*>(targets[i]);
    NF_REQUIRE(f.device == c.devices[i], "flows[%d] lives on device %d, the communicator expects device %d", i, f.device, c.devices[i]);
    NF_TRY(check_target(f, tg[i]));
    int64_t b, e;
    nf::shard_range(N_total, c.n_ranks, c.ranks[i], &b, &e);
    NF_CUDA(cudaSetDevice(f.device));
    f.ws_reset();
    NF_TRY(general_plan_workspace(f, OP_ELBO, e - b, 0));
  }
  return nf::multi_core(c, fl.data(), tg.data(), OP_ELBO, theta_dev, N_total, z0_dev, seed, scale, value_out, grad_dev_out);
}

}  // extern "C"
