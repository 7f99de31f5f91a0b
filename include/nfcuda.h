/*
 * libnfcuda -- C ABI of the B200-native (sm_100a) training hot path of NormalizingFlows.jl.
 *
 * This is the drop-in boundary described in SURVEY.md section 8(b).  A Julia package extension
 * binds these symbols with `ccall` (see INTEGRATION.md and julia/NormalizingFlowsNFCUDAExt.jl) and
 * plugs them in behind the reference's own seam
 *     _prepare_gradient / _value_and_gradient        (reference src/optimize.jl:8-14)
 * so that `train_flow(elbo, flow, logp, n; ADbackend=AutoNFCUDA())`
 *                                                      (reference src/NormalizingFlows.jl:54-86)
 * runs on the GPU with `optimize`, Optimisers.jl and the destructure'd theta untouched.
 *
 * Conventions
 *   - every function returns 0 on success and a negative nf_status on failure; the message is
 *     available from nf_last_error() (thread local).  Nothing throws or aborts across the ABI.
 *   - the caller owns every host buffer; the library owns all device memory behind opaque handles.
 *   - batches are Julia `d x N` column-major arrays == row-major [N][d]: each sample's d numbers are
 *     contiguous.  theta / grad are flat vectors in `Optimisers.destructure` order (reference
 *     src/NormalizingFlows.jl:67) with element type = the flow's dtype.
 *   - layers are listed in theta order, i.e. the order of `Ls` handed to `create_flow(Ls, q0)`
 *     (reference src/flows/utils.jl:23-26); the transform applies Ls[end] first.
 *   - mask indices are 0-based here (Julia side subtracts 1).
 *   - `scale` multiplies the returned value and gradient: pass -1.0 to obtain the loss
 *     `-vo(rng, re(theta), args...)` of reference src/NormalizingFlows.jl:69 and its gradient.
 *   - `*_dev` variants take device pointers (same layout) on the flow's device and enqueue on the
 *     flow's stream; they are what a multi-GPU host uses before its gradient all-reduce.
 *   - there is no CPU fallback: every entry point fails with NF_ERR_CUDA when no sm_100 device is
 *     usable.
 */
#ifndef NFCUDA_H
#define NFCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NFCUDA_VERSION 100

#if defined(__GNUC__)
#define NF_API __attribute__((visibility("default")))
#else
#define NF_API
#endif

typedef struct nf_flow_s*   nf_flow_t;
typedef struct nf_target_s* nf_target_t;

typedef enum nf_status {
  NF_OK              =  0,
  NF_ERR_INVALID     = -1,   /* bad argument / unsupported configuration */
  NF_ERR_CUDA        = -2,   /* CUDA runtime / driver error, or no usable device */
  NF_ERR_OOM         = -3,   /* device allocation failed */
  NF_ERR_UNSUPPORTED = -4    /* valid request this build cannot run (e.g. dim too large for a kernel) */
} nf_status;

typedef enum nf_dtype { NF_F32 = 0, NF_F64 = 1 } nf_dtype;

/* Layer kinds.  PLANAR/RADIAL: Bijectors.PlanarLayer/RadialLayer built by planarflow/radialflow
 * (reference src/flows/planar_radial.jl:21-29,52-60).  AFFINE_COUPLING: reference
 * src/flows/realnvp.jl:33-110.  SPLINE_COUPLING: reference src/flows/neuralspline.jl:35-140.
 * SHIFT/SCALE: Bijectors.Shift/Scale, the flow of the reference's analytic tests
 * (reference test/objectives.jl:9, test/interface.jl:22-24). */
typedef enum nf_layer_kind {
  NF_PLANAR          = 1,   /* theta: w(d), u(d), b(1)        */
  NF_RADIAL          = 2,   /* theta: alpha_(1), beta(1), z0(d) */
  NF_AFFINE_COUPLING = 3,   /* theta: s-chain, t-chain; chain = W1(:),b1,W2(:),b2,...  (W out x in, column major) */
  NF_SPLINE_COUPLING = 4,   /* theta: nn-chain with (3K-1)*n_mask outputs */
  NF_SHIFT           = 5,   /* theta: a(d) */
  NF_SCALE           = 6,   /* theta: a(d) */
  /* Hamiltonian flow on z = [x, rho], dim = 2h (reference example/demo_hamiltonian_flow.jl).  h a power of two <= 32 (<= 8 with the
   * logistic-regression score): fused one-thread-per-sample kernels, every entry point.  Any h <= 128 (BASELINE config 5: h = 100):
   * warp-per-sample kernels -- flows of Shift / Scale / these two kinds, LogReg / Funnel / Banana / DiagNormal scores, joint target;
   * both objectives (value + gradient), terms, forward, inverse, logpdf, sampling: */
  NF_MOMENTUM_AFFINE = 7,   /* Stacked((identity, Shift(b) ∘ Scale(a)), [1:h, h+1:2h]) (:94-99); theta: b(h), a(h) */
  NF_LEAPFROG        = 8    /* LeapFrog bijector (:27-91), logdet 0; theta: log_eps(h) (`@functor LeapFrog (logϵ,)` :39) */
} nf_layer_kind;

typedef struct nf_layer_desc {
  int        kind;       /* nf_layer_kind */
  const int* mask_idx;   /* couplings: 0-based indices of the TRANSFORMED coordinates (PartitionMask idx) */
  int        n_mask;
  const int* hdims;      /* couplings: hidden widths of the conditioner MLP(s) (`hdims` of fnn, reference src/flows/utils.jl:71-100) */
  int        n_hidden;
  int        K;          /* spline: number of bins   (reference src/flows/neuralspline.jl:37) */
  double     B;          /* spline: domain half-width (reference src/flows/neuralspline.jl:39) */
  int        n_steps;    /* leapfrog: number of leapfrog steps L (demo_hamiltonian_flow.jl:30) */
  const void* score_target; /* leapfrog: nf_target_t over the h position coordinates whose score drives the dynamics
                              (`∇logp`, :31); Banana, Funnel or DiagNormal; copied at nf_flow_create */
} nf_layer_desc;

/* Built-in target log-densities with device-side logp and score (reference example/targets/*.jl). */
typedef enum nf_target_kind {
  NF_TARGET_BANANA       = 1,  /* params: b, var            (banana.jl:77-83)           */
  NF_TARGET_FUNNEL       = 2,  /* params: mu, sigma         (neal_funnel.jl:54-61)      */
  NF_TARGET_WARPED_GAUSS = 3,  /* params: sigma1, sigma2    (warped_gaussian.jl:81-87)  */
  NF_TARGET_CROSS        = 4,  /* params: mu, sigma; dim = 2m -> product of m Cross blocks (cross.jl:30-38) */
  NF_TARGET_DIAG_NORMAL  = 5,  /* params: mu[dim], sigma[dim] (standard deviations; MvNormal(mu, Diagonal(sigma.^2))) */
  NF_TARGET_LOGREG       = 6   /* synthetic Bayesian logistic-regression posterior (BASELINE config 5): params: sigma0, n, X[n*dim] row major,
                                  y[n] in {0,1};  logp(b) = sum_i [y_i x_i.b - softplus(x_i.b)] + log N(b; 0, sigma0^2 I) */
} nf_target_kind;

/* MMA issue mode of the coupling-MLP contractions (fp32 flows).  F16X3 is the parity mode. */
typedef enum nf_mma_mode {
  NF_MMA_SIMT    = 0,  /* CUDA-core FMA GEMMs in the flow dtype (the only mode for NF_F64)            */
  NF_MMA_F16X3   = 1,  /* tcgen05 kind::f16: operands split hi+lo fp16 under an exact per-tensor power-of-two
                          scale (22 significant bits), 3 products, fp32 accumulate in TMEM                */
  NF_MMA_F16X1   = 2   /* tcgen05 single fp16 pass (11-bit operands): NOT parity-grade, throughput studies   */
} nf_mma_mode;

/* ---- library / device ------------------------------------------------------------------------ */
NF_API int         nf_version(void);
NF_API const char* nf_last_error(void);
/* Select the CUDA device this thread's subsequent handles live on; checks for compute capability 10.x. */
NF_API int         nf_init(int device);
NF_API int         nf_device_count(int* count);
NF_API int         nf_synchronize(void);

/* ---- flow ------------------------------------------------------------------------------------ */
/* Replaces: create_flow / planarflow / radialflow / realnvp / nsf structure + Optimisers.destructure layout
 * (reference src/flows/utils.jl:23-26, src/NormalizingFlows.jl:67). */
NF_API int     nf_flow_create(nf_flow_t* out, const nf_layer_desc* layers, int n_layers, int dim, int dtype);
NF_API void    nf_flow_destroy(nf_flow_t flow);
NF_API int64_t nf_flow_num_params(nf_flow_t flow);
NF_API int     nf_flow_dim(nf_flow_t flow);
/* Base distribution q0 = MvNormal(mu, Diagonal(sigma.^2)); NULL -> zeros / ones.  double arrays of length dim. */
NF_API int     nf_flow_set_base(nf_flow_t flow, const double* mu, const double* sigma);
/* Full-covariance base q0 = MvNormal(mu, Sigma), Sigma = L L^T (what reference ext/NormalizingFlowsCUDAExt.jl:43-47 samples
 * with `unwhiten!`; logpdf per Distributions).  L: dim x dim ROW-major lower-triangular Cholesky factor with
 * a positive diagonal (Julia: `permutedims(cholesky(Sigma).L)` of the column-major array, or pass `cholesky(Sigma).U`'s
 * memory as is); entries above the diagonal are ignored.  mu may be NULL (zeros).  Sampling, ELBO, log-likelihood, logpdf
 * and rand use it on every flow family (Hamiltonian flows: forward direction only).  nf_flow_set_base switches back to the
 * diagonal form. */
NF_API int     nf_flow_set_base_chol(nf_flow_t flow, const double* mu, const double* L);
/* nf_mma_mode for the coupling MLPs; default NF_MMA_F16X3 for NF_F32 flows, NF_MMA_SIMT for NF_F64. */
NF_API int     nf_flow_set_mma_mode(nf_flow_t flow, int mode);
/* Cap (bytes) on the activation workspace; batches larger than fits are processed in sample chunks. */
NF_API int     nf_flow_set_workspace_limit(nf_flow_t flow, size_t bytes);
/* Offset (in elements) of layer `layer`'s parameters inside theta. */
NF_API int64_t nf_flow_param_offset(nf_flow_t flow, int layer);

/* ---- targets --------------------------------------------------------------------------------- */
NF_API int  nf_target_create(nf_target_t* out, int kind, int dim, const double* params, int n_params);
/* logp_joint(z) = logp(x) + sum(logpdf(Normal(), rho)) on z = [x, rho] (reference example/demo_hamiltonian_flow.jl:117-124);
 * dimension 2 * dim(inner).  `inner` is copied. */
NF_API int  nf_target_create_joint(nf_target_t* out, nf_target_t inner);
NF_API void nf_target_destroy(nf_target_t target);
/* log-density and score of a device target at host-supplied points (x: [N][dim] of dtype; logp_out: [N]; score_out: [N][dim]
 * or NULL).  A host can check with it that the `logp` closure it was handed IS the device target it named (the Julia shim
 * does, before training), and tests pin the targets against the reference formulas (example/targets/*.jl). */
NF_API int nf_target_logp(nf_target_t target, int dtype, const void* x_host, int64_t N, void* logp_host_out, void* score_host_out);

/* ---- objectives: value and gradient ---------------------------------------------------------- */
/* Replaces _value_and_gradient(loss, prep, ad, theta, rng, logp, n) for vo = elbo / elbo_batch
 * (reference src/optimize.jl:12-14,86; src/objectives/elbo.jl:4-7,31-34,65-70,89-97).
 * z0_host: N x dim base draws (NULL -> drawn on device with Philox keyed by `seed`, replacing
 * _device_specific_rand, reference src/NormalizingFlows.jl:100-127, ext/NormalizingFlowsCUDAExt.jl:7-48).
 * value_out: scale * mean_j elbo_j.  grad_host_out: scale * d/dtheta (P elements of the flow dtype; may be NULL). */
NF_API int nf_elbo_value_and_grad(nf_flow_t flow, nf_target_t target, const void* theta_host, int64_t N,
                           const void* z0_host, uint64_t seed, double scale,
                           double* value_out, void* grad_host_out);
NF_API int nf_elbo_value_and_grad_dev(nf_flow_t flow, nf_target_t target, const void* theta_dev, int64_t N,
                               const void* z0_dev, uint64_t seed, double scale,
                               double* value_out, void* grad_dev_out);
/* Un-normalised sums for sample-sharded data parallelism: writes the SUM over the N local samples of
 * elbo_j into sums_dev_out[P] and of d elbo_j/dtheta into sums_dev_out[0..P) (P+1 elements of the flow
 * dtype, device memory) so that one all-reduce of P+1 numbers followed by a division by N_total
 * finishes the step. */
NF_API int nf_elbo_sums_dev(nf_flow_t flow, nf_target_t target, const void* theta_dev, int64_t N,
                     const void* z0_dev, uint64_t seed, void* sums_dev_out);
/* Per-sample ELBO terms elbo_j = logp(T(x_j)) - log q0(x_j) + logdet_j (reference elbo.jl:65-70). */
NF_API int nf_elbo_terms(nf_flow_t flow, nf_target_t target, const void* theta_host, int64_t N,
                  const void* z0_host, void* elbos_host_out);

/* Replaces _value_and_gradient for vo = loglikelihood (reference src/objectives/loglikelihood.jl:26-33):
 * value = scale * mean_j logpdf(flow, xs_j), gradient through the inverse flow. */
NF_API int nf_loglik_value_and_grad(nf_flow_t flow, const void* theta_host, int64_t N, const void* xs_host,
                             double scale, double* value_out, void* grad_host_out);
NF_API int nf_loglik_value_and_grad_dev(nf_flow_t flow, const void* theta_dev, int64_t N, const void* xs_dev,
                                 double scale, double* value_out, void* grad_dev_out);

/* On-device optimiser loop: `n_iters` iterations of
 *     ls, g = value_and_gradient(-elbo) ; theta <- Optimisers.Adam(eta, (beta1, beta2), eps) step
 * i.e. the body of the reference's `while` loop (src/optimize.jl:85-105) without its per-iteration host round trip;
 * valid when no callback / hasconverged needs theta on the host every iteration.  Base draws come from the device
 * Philox stream with seed + i for iteration i.  theta (and optionally the Adam moments m, v, for continuing a run that
 * already made t0 steps) are read from and written back to the host buffers once.  stats_out[i] = {loss_i, gradient_norm_i}
 * (the reference's opt_stats fields, src/optimize.jl:89). */
NF_API int nf_train_elbo_adam(nf_flow_t flow, nf_target_t target, void* theta_host_inout, int64_t N, uint64_t seed, int n_iters,
                              int t0, double eta, double beta1, double beta2, double eps, void* m_host_inout, void* v_host_inout,
                              double* stats_out);

/* ---- transform / density API ------------------------------------------------------------------ */
/* with_logabsdet_jacobian(flow.transform, xs): y_out N x dim, logdet_out N (either may be NULL).
 * (reference src/flows/realnvp.jl:77-83, src/flows/neuralspline.jl:102-108; Bijectors planar/radial) */
NF_API int nf_forward(nf_flow_t flow, const void* theta_host, int64_t N, const void* x_host,
               void* y_host_out, void* logdet_host_out);
/* with_logabsdet_jacobian(inverse(flow.transform), ys) (reference realnvp.jl:99-110, neuralspline.jl:133-140). */
NF_API int nf_inverse(nf_flow_t flow, const void* theta_host, int64_t N, const void* y_host,
               void* x_host_out, void* logdet_host_out);
/* logpdf(flow, ys) (Bijectors TransformedDistribution; used by reference test/flow.jl:15-16). */
NF_API int nf_logpdf(nf_flow_t flow, const void* theta_host, int64_t N, const void* y_host, void* logpdf_host_out);
/* rand(flow, N): base draws on device (Philox, `seed`) pushed through the flow in one batched pass;
 * replaces the per-column loop of reference ext/NormalizingFlowsCUDAExt.jl:65-74. */
NF_API int nf_sample(nf_flow_t flow, const void* theta_host, int64_t N, uint64_t seed, void* y_host_out);
/* randn draws N x dim ~ q0 (replaces _device_specific_rand for MvNormal, reference ext:41-48). */
NF_API int nf_base_sample(nf_flow_t flow, int64_t N, uint64_t seed, void* z_host_out);

/* Two-phase API for an arbitrary user log-density evaluated by the caller (SURVEY section 7 'Arbitrary Julia logp'):
 * nf_forward_stash runs the flow and keeps the activations; the caller evaluates logp(y) and
 * d logp/dy; nf_backward returns  sum_j [ gy_j . dy_j/dtheta + gld_j * dlogdet_j/dtheta ]. */
NF_API int nf_forward_stash(nf_flow_t flow, const void* theta_host, int64_t N, const void* x_host,
                     void* y_host_out, void* logdet_host_out);
NF_API int nf_backward(nf_flow_t flow, const void* gy_host, const void* gld_host_or_null, void* grad_host_out);

/* ---- test hooks (exercised by tests/, not by the Julia shim) ---------------------------------- */
/* Spline bin indices of every spline coupling for a forward pass over x_host: bins_out is
 * [n_spline_layers][N][n_mask] int32, in application order.  Bit-exactness check of BASELINE north_star. */
NF_API int nf_spline_bins(nf_flow_t flow, const void* theta_host, int64_t N, const void* x_host, int32_t* bins_out);
/* Bin search alone on caller-supplied knots: knots [M][K+1], v [M] -> bins [M] (searchsortedfirst - 1). */
NF_API int nf_rqs_bin_search(int dtype, const void* knots_host, const void* v_host, int64_t M, int K, int32_t* bins_out);
/* One Dense layer Y[n,N] = X[n,K] Wt[K,N] + b through the tcgen05 forward GEMM (terms: 1 or 3 fp16 products). */
NF_API int nf_tc_gemm_test(int64_t n, int K, int N, const float* X, const float* Wt, const float* b, int terms, float* Y);
/* Host-side test hook (no device needed): the per-network item schedule of the two-team fused coupling kernel for a hidden
 * width of n_chunks x 64 columns, accumulation chains of `chain` K chunks, n_hoist first-Dense chunks issued in the previous
 * network's tail and third-Dense slabs `delay` items after their chunk's last slab.  Item bytes: see csrc/fused_coupling.cuh.
 * Returns the number of items (written up to cap) or a negative error code. */
NF_API int nf_fused_schedule(int n_chunks, int chain, int n_hoist, int delay, unsigned char* items_out, int cap);
/* Number of kernels launched by this thread's library calls since the last reset (bench.py `gpu_launches`). */
NF_API int64_t nf_launch_count(int reset);
/* ---- multi-GPU data parallelism (the call that replaces reference src/optimize.jl:86 on G devices) ----------
 * Samples shard across devices in contiguous blocks (rank r owns rows [r*N/R, (r+1)*N/R), the remainder spread over the
 * first ranks), theta is replicated, and the only exchange is ONE ncclAllReduce(sum) of the P+1 double accumulators
 * (parameter-gradient sums + objective sum) on each device's compute stream, before the 1/N_total scaling.  NCCL is
 * resolved at run time (libnccl.so.2); nothing else in the library needs it.
 *   nf_comm_init_all : ONE process drives n_dev GPUs (the Julia host): ncclCommInitAll over dev_ids (NULL: 0..n_dev-1)
 *   nf_comm_unique_id / nf_comm_init_rank : one process per GPU (torchrun, MPI): rank 0 makes the id, the launcher hands
 *                      its NF_UNIQUE_ID_BYTES bytes to every rank, each rank joins with its own device                    */
typedef struct nf_comm_s* nf_comm_t;
#define NF_UNIQUE_ID_BYTES 128
NF_API int nf_comm_init_all(nf_comm_t* out, int n_dev, const int* dev_ids);
NF_API int nf_comm_unique_id(void* id_out);
NF_API int nf_comm_init_rank(nf_comm_t* out, int n_ranks, int rank, const void* id, int device);
NF_API int nf_comm_size(nf_comm_t comm);          /* ranks in the job                       */
NF_API int nf_comm_local_size(nf_comm_t comm);    /* devices driven by this process         */
NF_API int nf_comm_local_device(nf_comm_t comm, int i);
NF_API int nf_comm_local_rank(nf_comm_t comm, int i);
NF_API void nf_comm_destroy(nf_comm_t comm);
/* flows[i] / targets[i]: one replica per LOCAL device i, created while that device was current (nf_init(device)).
 * theta_host: P values.  N_total: samples over the whole job.  z0_host: the rows owned by THIS process (all of them with
 * nf_comm_init_all; the rank's own shard with one process per GPU), or NULL for device Philox draws -- rank r then draws
 * rows [begin_r, end_r) of the one global draw matrix of (seed), so the result equals the single-device result up to
 * summation order.  value_out / grad_host_out: the job-wide ELBO and gradient (identical on every process).            */
NF_API int nf_elbo_value_and_grad_multi(nf_comm_t comm, const nf_flow_t* flows, const nf_target_t* targets, const void* theta_host,
                                        int64_t N_total, const void* z0_host, uint64_t seed, double scale, double* value_out,
                                        void* grad_host_out);
/* forward-KL twin (reference src/objectives/loglikelihood.jl:26-33): xs_host = this process's rows of the data batch */
NF_API int nf_loglik_value_and_grad_multi(nf_comm_t comm, const nf_flow_t* flows, const void* theta_host, int64_t N_total,
                                          const void* xs_host, double scale, double* value_out, void* grad_host_out);
/* Same with everything resident: theta_dev[i] (P values), in_dev[i] (device i's shard, or NULL / null entries for Philox
 * draws), grad_dev_out[i] (P values, may be NULL) per local device; nothing crosses PCIe except the objective value.   */
NF_API int nf_elbo_value_and_grad_multi_dev(nf_comm_t comm, const nf_flow_t* flows, const nf_target_t* targets,
                                            const void* const* theta_dev, int64_t N_total, const void* const* z0_dev, uint64_t seed,
                                            double scale, double* value_out, void* const* grad_dev_out);
/* rows [begin, end) of a batch of N_total owned by `rank` of `n_ranks` (the partition the calls above use) */
NF_API void nf_shard_range(int64_t N_total, int n_ranks, int rank, int64_t* begin, int64_t* end);

/* Process-wide execution options (diagnostics / A-B measurements; results are parity grade either way).
 *   "fused_coupling"  1 (default): AffineCoupling layers that qualify run the fused conditioner kernels; 0: layer by layer
 *   "fused_variant"   0 (default): two-team streaming kernel (csrc/fused_coupling.cuh); 1: 128-column-MMA kernel
 *   "hmc_warp"        1: every Hamiltonian flow the warp-per-sample kernels cover runs on them (csrc/hmc_warp.cu; default 0:
 *                     only flows the one-thread-per-sample kernels cannot take -- h not a power of two, h > 32, or h > 8 with
 *                     the logistic-regression score)
 *   "rqs_planes"      1 (default): the spline-coupling backward kernel writes the gradient w.r.t. the conditioner output as
 *                     the split fp16 planes the GEMMs read, under a predicted power-of-two scale (exact-scale redo as the safety
 *                     net); 0: as an fp32 matrix plus a split pass; 2: planes with a deliberately wrong prediction, so that the
 *                     redo pass runs (tests)
 * Returns NF_ERR_INVALID for an unknown name. */
NF_API int nf_set_option(const char* name, int value);
/* Duration (ms, CUDA events on the flow's stream) of the device work of the last value_and_grad call. */
NF_API double  nf_last_device_ms(nf_flow_t flow);

/* Per-kernel-class device timing (CUDA events on the flow's stream), used by bench.py for the roofline:
 * enable, run, then list the recorded classes (comma separated) and collect {launches, total ms} per class. */
NF_API int nf_profile_enable(nf_flow_t flow, int on);
NF_API int nf_profile_keys(nf_flow_t flow, char* buf, int buflen);
NF_API int nf_profile_collect(nf_flow_t flow, const char* key, int64_t* launches, double* total_ms);

#ifdef __cplusplus
}
#endif
#endif /* NFCUDA_H */
