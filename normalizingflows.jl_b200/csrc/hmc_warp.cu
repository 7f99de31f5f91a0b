// K1h -- warp-per-sample Hamiltonian flow (BASELINE config 5: LeapFrog + momentum-affine layers on a 100-D posterior).
//
// Reference: example/demo_hamiltonian_flow.jl:27-99 (LeapFrog bijector, momentum normalisation layer), :117-124 (joint
// target logp(x) + sum logN(rho; 0, 1)), :139-147 (flow = 15 x (momentum affine o LeapFrog) over q0 = Shift o Scale (MvNormal)).
//
// The fused elementwise kernel (elementwise_impl.cuh) keeps a sample's whole state in ONE thread's registers, which stops at
// h = 32 (8 with the logistic-regression score).  Here a WARP owns a sample: lane l holds coordinates 4l .. 4l + 3 of
// the position x and of the momentum rho (h <= 128), the data set of the logistic-regression posterior is walked 32 observations
// at a time (coalesced row loads, one transposing reduction per round so that each lane evaluates one row's transcendental), and the reverse sweep needs no stash: every elementary update is undone
// exactly while its adjoint is applied, the second-order term being a Hessian-vector product of the target (same algorithm as
// leapfrog_backward in elementwise_impl.cuh).  Per-parameter gradient sums: shared-memory atomics per CTA, one partial row per CTA,
// finished in double by hw_finalize_kernel (layer-table layout and chain rules identical to ew_prep_body / ew_finalize_body).
//
// Forward direction: ELBO value / gradient, per-sample terms, transform + logdet.  Inverse direction: transform, logdet, the
// log-density head (logpdf) and the gradient of the forward-KL objective (reference src/objectives/loglikelihood.jl:26-33).
#include "flow.hpp"
#include "targets.cuh"

namespace nf {
namespace {

constexpr int HW_NPL = 4;            // coordinates per lane: h <= 128
constexpr int HW_THREADS = 128;      // four samples in flight per CTA

template <typename T> struct HwArgs {
  const T* z0;          // [N, d] (ignored with EW_GEN_Z0)
  const T* table;       // [L, 4 + 2 d]
  const int* kinds;     // [L]
  const T* base;        // mu[d], sigma[d] or nullptr
  T base_c0;
  TargetParams<T> tp;   // inner target of the joint objective (dim h)
  TargetParams<T> sp;   // score target of the LeapFrog layers (dim h)
  T* y_out; T* ld_out; T* terms_out;
  T* gpart;             // [grid, L * 2 d]
  double* epart;        // [grid]
  int64_t N;
  int L, d, flags;
  uint64_t seed;
  int64_t row0;
};

enum : int { HW_GRAD = 1, HW_TARGET = 2, HW_WRITE_Y = 4, HW_WRITE_LD = 8, HW_WRITE_TERMS = 16, HW_GEN_Z0 = 32 };

// this lane's four entries of a data row (zero past h): one 16-byte (two for double) load when the rows are aligned
template <typename T>
__device__ __forceinline__ void hw_load_row(const T* __restrict__ xr, int h, int lane, bool vec, T (&xv)[HW_NPL]) {
  const int k0 = 4 * lane;
  if (vec) {
    if (k0 < h) {
      if constexpr (sizeof(T) == 4) {
        const float4 t = *reinterpret_cast<const float4*>(xr + k0);
        xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
      } else {
        const double2 t0 = *reinterpret_cast<const double2*>(xr + k0), t1 = *reinterpret_cast<const double2*>(xr + k0 + 2);
        xv[0] = t0.x; xv[1] = t0.y; xv[2] = t1.x; xv[3] = t1.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < HW_NPL; ++i) xv[i] = 0;
    }
  } else {
#pragma unroll
    for (int i = 0; i < HW_NPL; ++i) xv[i] = (k0 + i < h) ? xr[k0 + i] : T(0);
  }
}

// Lane l receives the sum over the 32 lanes of p[l % R] (p is consumed): recursive halving, 31 shuffles for R = 32 or 16
// instead of R x 5
template <typename T, int R>
__device__ __forceinline__ T hw_transpose_sum(T (&p)[R], int lane) {
  if constexpr (R == 16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) p[j] += __shfl_xor_sync(0xffffffffu, p[j], 16);
  }
#pragma unroll
  for (int s = (R == 32 ? 16 : 8); s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const T send = up ? p[j] : p[j + s];
      const T keep = up ? p[j + s] : p[j];
      p[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return p[0];
}
// observations per round: the rows of a round stay in registers between the dot products and the rank-one updates
template <typename T> constexpr int hw_round() { return sizeof(T) == 4 ? 32 : 16; }

// logp and score of the (inner) target at x; every lane returns the same logp
template <typename T>
__device__ __forceinline__ T hw_score(const TargetParams<T>& tp, const T (&x)[HW_NPL], T (&g)[HW_NPL], int lane) {
  using N_ = Num<T>;
  const int h = tp.dim;
  switch (tp.kind) {
    case NF_TARGET_LOGREG: {   // u_i = x_i . z ; logp = sum_i [y_i u_i - softplus(u_i)] - |z|^2 / (2 sigma0^2) + c0   (targets.cuh)
      // R observations per round (32 float / 16 double): every lane forms its share of the R dot products from vector row loads, a
      // transposing reduction hands observation r to lane r, which evaluates the ONE softplus / sigmoid of that row; the
      // residuals then go back by broadcast for the rank-one score updates on the rows still held in registers
      const T is2 = 1 / (tp.p0 * tp.p0);
      const T* X = tp.vec;
      const T* y = tp.vec + (size_t)tp.n_data * h;
      const bool vec = (h & 3) == 0;
      T q = 0, lp = 0;
#pragma unroll
      for (int i = 0; i < HW_NPL; ++i) { q += x[i] * x[i]; g[i] = -x[i] * is2; }
      constexpr int R = hw_round<T>();
      for (int r0 = 0; r0 < tp.n_data; r0 += R) {
        T part[R], xk[R][HW_NPL];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
          T u = 0;
#pragma unroll
          for (int i = 0; i < HW_NPL; ++i) xk[rr][i] = 0;
          if (r0 + rr < tp.n_data) {
            hw_load_row<T>(X + (size_t)(r0 + rr) * h, h, lane, vec, xk[rr]);
#pragma unroll
            for (int i = 0; i < HW_NPL; ++i) u += xk[rr][i] * x[i];
          }
          part[rr] = u;
        }
        const T u = hw_transpose_sum<T, R>(part, lane);
        const int row = r0 + (lane & (R - 1));
        T res = 0;
        if (row < tp.n_data) {
          const T yr = y[row];
          if (lane < R) lp += yr * u - softplus_stable<T>(u);
          res = yr - sigmoid_stable<T>(u);
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
          const T rv = __shfl_sync(0xffffffffu, res, rr);        // rows past n_data hold zeros and rv = 0
#pragma unroll
          for (int i = 0; i < HW_NPL; ++i) g[i] += rv * xk[rr][i];
        }
      }
      q = warp_sum(q); lp = warp_sum(lp);
      return tp.c0 + lp - q * is2 / 2;
    }
    case NF_TARGET_BANANA: {   // banana.jl:77-83: b = p0, var = p1 (coordinates 0 and 1 live in lane 0)
      const T b = tp.p0, vr = tp.p1;
      const T u1 = __shfl_sync(0xffffffffu, x[0], 0);
      const T u2 = __shfl_sync(0xffffffffu, x[1], 0) + b * u1 * u1 - vr * b;
      T q = 0;
#pragma unroll
      for (int i = 0; i < HW_NPL; ++i) {
        const int k = 4 * lane + i;
        g[i] = 0;
        if (k >= 2 && k < h) { q += x[i] * x[i]; g[i] = -x[i]; }
      }
      q = warp_sum(q) + u1 * u1 / vr + u2 * u2;
      if (lane == 0) { g[0] = -u1 / vr - u2 * (2 * b * u1); g[1] = -u2; }
      return tp.c0 - q / 2;
    }
    case NF_TARGET_FUNNEL: {   // neal_funnel.jl:54-72: mu = p0, sigma = p1
      const T mu = tp.p0, sg = tp.p1;
      const T x1 = __shfl_sync(0xffffffffu, x[0], 0);
      const T a = N_::exp(-x1);
      T ss = 0;
#pragma unroll
      for (int i = 0; i < HW_NPL; ++i) {
        const int k = 4 * lane + i;
        if (k >= 1 && k < h) { ss += x[i] * x[i]; g[i] = -a * x[i]; } else g[i] = 0;
      }
      ss = warp_sum(ss);
      if (lane == 0) g[0] = (mu - x1) / (sg * sg) - T(h - 1) / 2 + a * ss / 2;
      return tp.c0 - (x1 - mu) * (x1 - mu) / (2 * sg * sg) - T(h - 1) / 2 * x1 - a * ss / 2;
    }
    case NF_TARGET_DIAG_NORMAL: {
      T q = 0;
#pragma unroll
      for (int i = 0; i < HW_NPL; ++i) {
        const int k = 4 * lane + i;
        g[i] = 0;
        if (k < h) {
          const T is = 1 / tp.vec[h + k];
          const T u = (x[i] - tp.vec[k]) * is;
          q += u * u;
          g[i] = -u * is;
        }
      }
      q = warp_sum(q);
      return tp.c0 - q / 2;
    }
  }
  return 0;
}

// out = (Hessian of logp at x) w
template <typename T>
__device__ __forceinline__ void hw_hvp(const TargetParams<T>& tp, const T (&x)[HW_NPL], const T (&w)[HW_NPL], T (&out)[HW_NPL], int lane) {
  using N_ = Num<T>;
  const int h = tp.dim;
  switch (tp.kind) {
    case NF_TARGET_LOGREG: {   // H w = -X^T diag(s (1 - s)) X w - w / sigma0^2,  s = sigmoid(X x)
      const T is2 = 1 / (tp.p0 * tp.p0);
      const T* X = tp.vec;
      const bool vec = (h & 3) == 0;
#pragma unroll
      for (int i = 0; i < HW_NPL; ++i) out[i] = -w[i] * is2;
      constexpr int R = hw_round<T>();
      for (int r0 = 0; r0 < tp.n_data; r0 += R) {
        T pu[R], pw[R], xk[R][HW_NPL];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
          T u = 0, xw = 0;
#pragma unroll
          for (int i = 0; i < HW_NPL; ++i) xk[rr][i] = 0;
          if (r0 + rr < tp.n_data) {
            hw_load_row<T>(X + (size_t)(r0 + rr) * h, h, lane, vec, xk[rr]);
#pragma unroll
            for (int i = 0; i < HW_NPL; ++i) { u += xk[rr][i] * x[i]; xw += xk[rr][i] * w[i]; }
          }
          pu[rr] = u; pw[rr] = xw;
        }
        const T u = hw_transpose_sum<T, R>(pu, lane);
        const T xw = hw_transpose_sum<T, R>(pw, lane);
        T cfac = 0;
        if (r0 + (lane & (R - 1)) < tp.n_data) {
          const T sg = sigmoid_stable<T>(u);
          cfac = -sg * (1 - sg) * xw;
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
          const T cv = __shfl_sync(0xffffffffu, cfac, rr);
#pragma unroll
          for (int i = 0; i < HW_NPL; ++i) out[i] += cv * xk[rr][i];
        }
      }
      return;
    }
    case NF_TARGET_BANANA: {
      const T b = tp.p0, vr = tp.p1;
      const T x0 = __shfl_sync(0xffffffffu, x[0], 0), x1 = __shfl_sync(0xffffffffu, x[1], 0);
      const T u2 = x1 + b * x0 * x0 - vr * b;
      const T h11 = -1 / vr - 2 * b * u2 - 4 * b * b * x0 * x0, h12 = -2 * b * x0;
#pragma unroll
      for (int i = 0; i < HW_NPL; ++i) { const int k = 4 * lane + i; out[i] = (k >= 2 && k < h) ? -w[i] : T(0); }
      if (lane == 0) { const T w0 = w[0], w1 = w[1]; out[0] = h11 * w0 + h12 * w1; out[1] = h12 * w0 - w1; }
      return;
    }
    case NF_TARGET_FUNNEL: {
      const T sg = tp.p1;
      const T x1 = __shfl_sync(0xffffffffu, x[0], 0), w1 = __shfl_sync(0xffffffffu, w[0], 0);
      const T a = N_::exp(-x1);
      T ss = 0, xw = 0;
#pragma unroll
      for (int i = 0; i < HW_NPL; ++i) {
        const int k = 4 * lane + i;
        out[i] = 0;
        if (k >= 1 && k < h) { ss += x[i] * x[i]; xw += x[i] * w[i]; out[i] = a * (x[i] * w1 - w[i]); }
      }
      ss = warp_sum(ss); xw = warp_sum(xw);
      if (lane == 0) out[0] = (-1 / (sg * sg) - a * ss / 2) * w1 + a * xw;
      return;
    }
    case NF_TARGET_DIAG_NORMAL: {
#pragma unroll
      for (int i = 0; i < HW_NPL; ++i) {
        const int k = 4 * lane + i;
        out[i] = 0;
        if (k < h) { const T is = 1 / tp.vec[h + k]; out[i] = -w[i] * is * is; }
      }
      return;
    }
  }
}

//   rho += eps/2 .* s(x);  { x += eps .* rho;  rho += eps .* s(x) } x (L-1);  x += eps .* rho;  rho += eps/2 .* s(x)
template <typename T>
__device__ __forceinline__ void hw_leapfrog_apply(const TargetParams<T>& sp, T (&x)[HW_NPL], T (&v)[HW_NPL], const T (&eps)[HW_NPL],
                                                  int nsteps, int lane) {
  T g[HW_NPL];
  hw_score<T>(sp, x, g, lane);
#pragma unroll
  for (int i = 0; i < HW_NPL; ++i) v[i] += eps[i] / 2 * g[i];
  for (int it = 0; it < nsteps; ++it) {
#pragma unroll
    for (int i = 0; i < HW_NPL; ++i) x[i] += eps[i] * v[i];
    hw_score<T>(sp, x, g, lane);
    const T c = (it == nsteps - 1) ? T(1) / 2 : T(1);
#pragma unroll
    for (int i = 0; i < HW_NPL; ++i) v[i] += c * eps[i] * g[i];
  }
}

// (x, v): OUTPUT state on entry, input state on exit; (gx, gv): adjoints of the output on entry, of the input on exit
template <typename T>
__device__ __forceinline__ void hw_leapfrog_backward(const TargetParams<T>& sp, T (&x)[HW_NPL], T (&v)[HW_NPL], T (&gx)[HW_NPL],
                                                     T (&gv)[HW_NPL], const T (&eps)[HW_NPL], int nsteps, T (&geps)[HW_NPL], int lane) {
  T g[HW_NPL], w[HW_NPL], hw[HW_NPL];
  for (int it = nsteps - 1; it >= 0; --it) {
    const T c = (it == nsteps - 1) ? T(1) / 2 : T(1);
    hw_score<T>(sp, x, g, lane);
#pragma unroll
    for (int i = 0; i < HW_NPL; ++i) {
      v[i] -= c * eps[i] * g[i];
      geps[i] += c * gv[i] * g[i];
      w[i] = c * eps[i] * gv[i];
    }
    hw_hvp<T>(sp, x, w, hw, lane);
#pragma unroll
    for (int i = 0; i < HW_NPL; ++i) gx[i] += hw[i];
#pragma unroll
    for (int i = 0; i < HW_NPL; ++i) {
      x[i] -= eps[i] * v[i];
      geps[i] += gx[i] * v[i];
      gv[i] += eps[i] * gx[i];
    }
  }
  hw_score<T>(sp, x, g, lane);
#pragma unroll
  for (int i = 0; i < HW_NPL; ++i) {
    v[i] -= eps[i] / 2 * g[i];
    geps[i] += gv[i] * g[i] / 2;
    w[i] = eps[i] / 2 * gv[i];
  }
  hw_hvp<T>(sp, x, w, hw, lane);
#pragma unroll
  for (int i = 0; i < HW_NPL; ++i) gx[i] += hw[i];
}

// layer table, runtime stride 4 + 2 d (same entries as ew_prep_body)
template <typename T>
__global__ void hw_prep_kernel(const T* __restrict__ theta, const EwLayerMeta* __restrict__ meta, int L, int d, T* __restrict__ table) {
  using N_ = Num<T>;
  const int l = blockIdx.x;
  if (l >= L) return;
  const int str = 4 + 2 * d, h = d / 2;
  const T* p = theta + meta[l].theta_off;
  T* e = table + (size_t)l * str;
  T* v0 = e + 4;
  T* v1 = e + 4 + d;
  const int kind = meta[l].kind;
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    T a0 = 0, a1 = 0;
    switch (kind) {
      case NF_SHIFT: a0 = p[k]; break;
      case NF_SCALE: a0 = p[k]; a1 = 1 / p[k]; break;
      case NF_MOMENTUM_AFFINE: a0 = k >= h ? p[k - h] : T(0); a1 = k >= h ? p[k] : T(1); break;
      case NF_LEAPFROG: a0 = k < h ? N_::exp(p[k]) : T(0); break;
    }
    v0[k] = a0; v1[k] = a1;
  }
  if (threadIdx.x == 0) {
    T c0 = 0;
    if (kind == NF_SCALE) for (int k = 0; k < d; ++k) c0 += N_::log(N_::abs(p[k]));
    if (kind == NF_MOMENTUM_AFFINE) for (int k = 0; k < h; ++k) c0 += N_::log(N_::abs(p[h + k]));
    if (kind == NF_LEAPFROG) c0 = (T)meta[l].aux;
    e[0] = c0; e[1] = 0; e[2] = 0; e[3] = 0;
  }
}

template <typename T>
__global__ void __launch_bounds__(HW_THREADS) hw_flow_kernel(HwArgs<T> a) {
  extern __shared__ __align__(16) unsigned char hw_smem[];
  T* s_acc = reinterpret_cast<T*>(hw_smem);      // [L][2 d]
  __shared__ double s_e[HW_THREADS / 32];
  const int L = a.L, d = a.d, h = d / 2, str = 4 + 2 * d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool want_grad = a.flags & HW_GRAD;
  if (want_grad) for (int i = tid; i < L * 2 * d; i += HW_THREADS) s_acc[i] = 0;
  __syncthreads();
  double elbo_local = 0;
  const int64_t nw = (int64_t)gridDim.x * (HW_THREADS / 32);
  for (int64_t j = (int64_t)blockIdx.x * (HW_THREADS / 32) + warp; j < a.N; j += nw) {
    T x[HW_NPL], v[HW_NPL];
    T q = 0;
    // ---- base draws, base log-density ----
#pragma unroll
    for (int i = 0; i < HW_NPL; ++i) {
      const int k = 4 * lane + i;
      x[i] = 0; v[i] = 0;
      if (k < h) {
        T zx, zv;
        if (a.flags & HW_GEN_Z0) {
          zx = philox_randn<T>(a.seed, (uint64_t)(a.row0 + j) * d + k);
          zv = philox_randn<T>(a.seed, (uint64_t)(a.row0 + j) * d + h + k);
        } else {
          zx = a.z0[j * d + k]; zv = a.z0[j * d + h + k];
        }
        if (a.base) {
          if (a.flags & HW_GEN_Z0) {
            q += zx * zx + zv * zv;
            zx = zx * a.base[d + k] + a.base[k];
            zv = zv * a.base[d + h + k] + a.base[h + k];
          } else {
            const T ux = (zx - a.base[k]) / a.base[d + k], uv = (zv - a.base[h + k]) / a.base[d + h + k];
            q += ux * ux + uv * uv;
          }
        } else {
          q += zx * zx + zv * zv;
        }
        x[i] = zx; v[i] = zv;
      }
    }
    q = warp_sum(q);
    const T lq = a.base_c0 - q / 2;
    T ld = 0;
    // ---- forward sweep: layers applied last-to-first (create_flow, reference src/flows/utils.jl:23-26) ----
    for (int l = L - 1; l >= 0; --l) {
      const T* e = a.table + (size_t)l * str;
      const T* v0 = e + 4;
      const T* v1 = e + 4 + d;
      const int kind = a.kinds[l];
      if (kind == NF_LEAPFROG) {
        T eps[HW_NPL];
#pragma unroll
        for (int i = 0; i < HW_NPL; ++i) { const int k = 4 * lane + i; eps[i] = k < h ? v0[k] : T(0); }
        hw_leapfrog_apply<T>(a.sp, x, v, eps, (int)e[0], lane);
        continue;
      }
#pragma unroll
      for (int i = 0; i < HW_NPL; ++i) {
        const int k = 4 * lane + i;
        if (k >= h) continue;
        if (kind == NF_SHIFT) { x[i] += v0[k]; v[i] += v0[h + k]; }
        else if (kind == NF_SCALE) { x[i] *= v0[k]; v[i] *= v0[h + k]; }
        else v[i] = v[i] * v1[h + k] + v0[h + k];      // NF_MOMENTUM_AFFINE
      }
      if (kind != NF_SHIFT) ld += e[0];
    }
    if (a.flags & HW_WRITE_Y) {
#pragma unroll
      for (int i = 0; i < HW_NPL; ++i) {
        const int k = 4 * lane + i;
        if (k < h) { a.y_out[j * d + k] = x[i]; a.y_out[j * d + h + k] = v[i]; }
      }
    }
    if ((a.flags & HW_WRITE_LD) && lane == 0) a.ld_out[j] = ld;
    if (!(a.flags & HW_TARGET)) continue;
    // ---- joint target: logp(x) + sum logN(rho; 0, 1)   (demo_hamiltonian_flow.jl:117-124) ----
    T gx[HW_NPL], gv[HW_NPL];
    T lp = hw_score<T>(a.tp, x, gx, lane);
    T qv = 0;
#pragma unroll
    for (int i = 0; i < HW_NPL; ++i) { qv += v[i] * v[i]; gv[i] = -v[i]; }
    qv = warp_sum(qv);
    lp -= qv / 2 + T(h) * T(NF_LOG2PI / 2);
    const T term = lp - lq + ld;
    if (lane == 0) {
      elbo_local += (double)term;
      if (a.flags & HW_WRITE_TERMS) a.terms_out[j] = term;
    }
    if (!want_grad) continue;
    // ---- backward sweep (reverse of application order); dELBO/dlogdet = 1 per sample ----
    for (int l = 0; l < L; ++l) {
      const T* e = a.table + (size_t)l * str;
      const T* v0 = e + 4;
      const T* v1 = e + 4 + d;
      T* acc = s_acc + (size_t)l * 2 * d;
      const int kind = a.kinds[l];
      if (kind == NF_LEAPFROG) {
        T eps[HW_NPL], ge[HW_NPL];
#pragma unroll
        for (int i = 0; i < HW_NPL; ++i) { const int k = 4 * lane + i; eps[i] = k < h ? v0[k] : T(0); ge[i] = 0; }
        hw_leapfrog_backward<T>(a.sp, x, v, gx, gv, eps, (int)e[0], ge, lane);
#pragma unroll
        for (int i = 0; i < HW_NPL; ++i) { const int k = 4 * lane + i; if (k < h) atomicAdd(&acc[k], ge[i]); }
        continue;
      }
#pragma unroll
      for (int i = 0; i < HW_NPL; ++i) {
        const int k = 4 * lane + i;
        if (k >= h) continue;
        if (kind == NF_SHIFT) {
          atomicAdd(&acc[k], gx[i]); atomicAdd(&acc[h + k], gv[i]);
          x[i] -= v0[k]; v[i] -= v0[h + k];
        } else if (kind == NF_SCALE) {
          x[i] *= v1[k]; v[i] *= v1[h + k];
          atomicAdd(&acc[k], gx[i] * x[i]); atomicAdd(&acc[h + k], gv[i] * v[i]);
          gx[i] *= v0[k]; gv[i] *= v0[h + k];
        } else {                                        // NF_MOMENTUM_AFFINE
          v[i] = (v[i] - v0[h + k]) / v1[h + k];
          atomicAdd(&acc[h + k], gv[i]); atomicAdd(&acc[d + h + k], gv[i] * v[i]);
          gv[i] *= v1[h + k];
        }
      }
    }
  }
  // ---- per-CTA partials ----
  __syncthreads();
  if (a.gpart && want_grad)
    for (int i = tid; i < L * 2 * d; i += HW_THREADS) a.gpart[(size_t)blockIdx.x * L * 2 * d + i] = s_acc[i];
  if (a.epart) {
    if (lane == 0) s_e[warp] = elbo_local;
    __syncthreads();
    if (tid == 0) {
      double t = 0;
      for (int w = 0; w < HW_THREADS / 32; ++w) t += s_e[w];
      a.epart[blockIdx.x] = t;
    }
  }
}

// Inverse direction: x = T^{-1}(y) (theta order: the first layer's inverse first), logdet of the inverse, and -- with HW_TARGET --
// the log-density head  log q(y) = logpdf(q0, x) + logdet_inv and its gradient (reference src/objectives/loglikelihood.jl:26-33).  LeapFrog^{-1} is the same map with -eps (demo_hamiltonian_flow.jl:63-82).
template <typename T>
__global__ void __launch_bounds__(HW_THREADS) hw_inv_kernel(HwArgs<T> a) {
  extern __shared__ __align__(16) unsigned char hw_smem[];
  T* s_acc = reinterpret_cast<T*>(hw_smem);      // [L][2 d] (HW_GRAD only)
  __shared__ double s_e[HW_THREADS / 32];
  const int L = a.L, d = a.d, h = d / 2, str = 4 + 2 * d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool want_grad = a.flags & HW_GRAD;
  if (want_grad) for (int i = tid; i < L * 2 * d; i += HW_THREADS) s_acc[i] = 0;
  __syncthreads();
  double obj_local = 0;
  const int64_t nw = (int64_t)gridDim.x * (HW_THREADS / 32);
  for (int64_t j = (int64_t)blockIdx.x * (HW_THREADS / 32) + warp; j < a.N; j += nw) {
    T x[HW_NPL], v[HW_NPL];
#pragma unroll
    for (int i = 0; i < HW_NPL; ++i) {
      const int k = 4 * lane + i;
      x[i] = k < h ? a.z0[j * d + k] : T(0);
      v[i] = k < h ? a.z0[j * d + h + k] : T(0);
    }
    T ld = 0;
    for (int l = 0; l < L; ++l) {
      const T* e = a.table + (size_t)l * str;
      const T* v0 = e + 4;
      const T* v1 = e + 4 + d;
      const int kind = a.kinds[l];
      if (kind == NF_LEAPFROG) {
        T eps[HW_NPL];
#pragma unroll
        for (int i = 0; i < HW_NPL; ++i) { const int k = 4 * lane + i; eps[i] = k < h ? -v0[k] : T(0); }
        hw_leapfrog_apply<T>(a.sp, x, v, eps, (int)e[0], lane);
        continue;
      }
#pragma unroll
      for (int i = 0; i < HW_NPL; ++i) {
        const int k = 4 * lane + i;
        if (k >= h) continue;
        if (kind == NF_SHIFT) { x[i] -= v0[k]; v[i] -= v0[h + k]; }
        else if (kind == NF_SCALE) { x[i] *= v1[k]; v[i] *= v1[h + k]; }
        else v[i] = (v[i] - v0[h + k]) / v1[h + k];      // NF_MOMENTUM_AFFINE
      }
      if (kind != NF_SHIFT) ld -= e[0];
    }
    if (a.flags & HW_WRITE_Y) {
#pragma unroll
      for (int i = 0; i < HW_NPL; ++i) {
        const int k = 4 * lane + i;
        if (k < h) { a.y_out[j * d + k] = x[i]; a.y_out[j * d + h + k] = v[i]; }
      }
    }
    if ((a.flags & HW_WRITE_LD) && lane == 0) a.ld_out[j] = ld;
    if (!(a.flags & HW_TARGET)) continue;
    T q = 0;
    T gx[HW_NPL], gv[HW_NPL];                      // d log q0 / d x0
#pragma unroll
    for (int i = 0; i < HW_NPL; ++i) {
      const int k = 4 * lane + i;
      gx[i] = 0; gv[i] = 0;
      if (k >= h) continue;
      T ux = x[i], uv = v[i], isx = 1, isv = 1;
      if (a.base) {
        isx = 1 / a.base[d + k]; isv = 1 / a.base[d + h + k];
        ux = (ux - a.base[k]) * isx; uv = (uv - a.base[h + k]) * isv;
      }
      q += ux * ux + uv * uv;
      gx[i] = -ux * isx; gv[i] = -uv * isv;
    }
    q = warp_sum(q);
    const T term = a.base_c0 - q / 2 + ld;
    if (lane == 0) {
      obj_local += (double)term;
      if (a.flags & HW_WRITE_TERMS) a.terms_out[j] = term;
    }
    if (!want_grad) continue;
    // ---- backward sweep: layers L-1 .. 0, walking the forward maps from x0 back to y (ew_inv_flow_kernel's scheme) ----
    for (int l = L - 1; l >= 0; --l) {
      const T* e = a.table + (size_t)l * str;
      const T* v0 = e + 4;
      const T* v1 = e + 4 + d;
      T* acc = s_acc + (size_t)l * 2 * d;
      const int kind = a.kinds[l];
      if (kind == NF_LEAPFROG) {                    // the inverse applied the map with -eps: d/d(eps) = -d/d(-eps)
        T eps[HW_NPL], ge[HW_NPL];
#pragma unroll
        for (int i = 0; i < HW_NPL; ++i) { const int k = 4 * lane + i; eps[i] = k < h ? -v0[k] : T(0); ge[i] = 0; }
        hw_leapfrog_backward<T>(a.sp, x, v, gx, gv, eps, (int)e[0], ge, lane);
#pragma unroll
        for (int i = 0; i < HW_NPL; ++i) { const int k = 4 * lane + i; if (k < h) atomicAdd(&acc[k], -ge[i]); }
        continue;
      }
#pragma unroll
      for (int i = 0; i < HW_NPL; ++i) {
        const int k = 4 * lane + i;
        if (k >= h) continue;
        if (kind == NF_SHIFT) {                     // z = y - a
          atomicAdd(&acc[k], -gx[i]); atomicAdd(&acc[h + k], -gv[i]);
          x[i] += v0[k]; v[i] += v0[h + k];
        } else if (kind == NF_SCALE) {              // z = y / a
          atomicAdd(&acc[k], -gx[i] * x[i] * v1[k]); atomicAdd(&acc[h + k], -gv[i] * v[i] * v1[h + k]);
          gx[i] *= v1[k]; gv[i] *= v1[h + k];
          x[i] *= v0[k]; v[i] *= v0[h + k];
        } else {                                    // NF_MOMENTUM_AFFINE: rho_in = (rho_out - b) / a
          gv[i] /= v1[h + k];
          atomicAdd(&acc[h + k], -gv[i]); atomicAdd(&acc[d + h + k], -gv[i] * v[i]);
          v[i] = v[i] * v1[h + k] + v0[h + k];
        }
      }
    }
  }
  __syncthreads();
  if (a.gpart && want_grad)
    for (int i = tid; i < L * 2 * d; i += HW_THREADS) a.gpart[(size_t)blockIdx.x * L * 2 * d + i] = s_acc[i];
  if (a.epart) {
    if (lane == 0) s_e[warp] = obj_local;
    __syncthreads();
    if (tid == 0) {
      double t = 0;
      for (int w = 0; w < HW_THREADS / 32; ++w) t += s_e[w];
      a.epart[blockIdx.x] = t;
    }
  }
}

// chain rule to theta (ew_finalize_body with a runtime stride): one CTA per layer
template <typename T>
__global__ void hw_finalize_kernel(const T* __restrict__ theta, const EwLayerMeta* __restrict__ meta, int L, int d,
                                   const T* __restrict__ gpart, const double* __restrict__ epart, int nblocks, int64_t N, int64_t P,
                                   int want_grad, int inverse, double* __restrict__ gsum) {
  const int l = blockIdx.x, h = d / 2;
  if (l == 0 && threadIdx.x == 0 && epart) {
    double e = 0;
    for (int b = 0; b < nblocks; ++b) e += epart[b];
    gsum[P] = e;
  }
  if (l >= L || !want_grad) return;
  const T* p = theta + meta[l].theta_off;
  double* g = gsum + meta[l].theta_off;
  const int kind = meta[l].kind;
  auto G = [&](int i) {
    double s = 0;
    for (int b = 0; b < nblocks; ++b) s += (double)gpart[((size_t)b * L + l) * 2 * d + i];
    return s;
  };
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    switch (kind) {
      case NF_SHIFT: g[k] = G(k); break;
      case NF_SCALE: g[k] = G(k) + (inverse ? -1.0 : 1.0) * (double)N / (double)p[k]; break;
      case NF_MOMENTUM_AFFINE:
        if (k < h) g[k] = G(h + k);
        else g[k] = G(d + k) + (inverse ? -1.0 : 1.0) * (double)N / (double)p[k];
        break;
      case NF_LEAPFROG:
        if (k < h) g[k] = G(k) * exp((double)p[k]);
        break;
    }
  }
}

bool hw_target_ok(int kind) { return kind == NF_TARGET_LOGREG || kind == NF_TARGET_FUNNEL || kind == NF_TARGET_DIAG_NORMAL || kind == NF_TARGET_BANANA; }

}  // namespace

// Does this flow / objective run on the warp-per-sample path?  (forward direction only)
bool hmc_warp_qualifies(const Flow& f, const Target* tgt) {
  if (!f.hamiltonian || f.base_dense || (f.dim & 1) || f.dim / 2 > 32 * HW_NPL) return false;
  for (const LayerDesc& L : f.layers)
    if (L.kind != NF_SHIFT && L.kind != NF_SCALE && L.kind != NF_MOMENTUM_AFFINE && L.kind != NF_LEAPFROG) return false;
  if (f.score_target && !hw_target_ok(f.score_target->kind)) return false;
  if (tgt && !(tgt->joint && hw_target_ok(tgt->kind))) return false;
  return true;
}

// workspace of one call (layer table, per-CTA gradient partials, per-CTA objective partials), either precision
size_t hmc_warp_workspace_bytes(const Flow& f) {
  if (!f.hamiltonian) return 0;
  const size_t L = f.layers.size(), d = (size_t)f.dim;
  return (L * (4 + 2 * d) + (size_t)kNumSMs * 4 * L * 2 * d + (size_t)kNumSMs * 4) * 8 + 4096;
}

template <typename T>
int hmc_warp_run(Flow& f, const Target* tgt, const void* theta_dev, int64_t N, const void* z0_dev, uint64_t seed, bool want_grad,
                 void* y_out, void* ld_out, void* terms_out, double* gsum_dev) {
  const int L = (int)f.layers.size(), d = f.dim;
  const size_t smem = want_grad ? (size_t)L * 2 * d * sizeof(T) : 16;
  if (smem > 200 * 1024) {
    set_error("Hamiltonian flow with %d layers over dim %d needs %zu B of shared memory for its gradient sums (limit 200 KiB)", L, d, smem);
    return NF_ERR_UNSUPPORTED;
  }
  NF_CUDA(cudaFuncSetAttribute(hw_flow_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int max_blocks = 0;
  NF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, hw_flow_kernel<T>, HW_THREADS, smem));
  if (max_blocks < 1) max_blocks = 1;
  const int grid = (int)std::min<int64_t>(ceil_div(N, (int64_t)(HW_THREADS / 32)), (int64_t)kNumSMs * std::min(max_blocks, 4));
  T* table = (T*)f.ws_alloc((size_t)L * (4 + 2 * d) * sizeof(T));
  T* gpart = (T*)f.ws_alloc((size_t)grid * L * 2 * d * sizeof(T));
  double* epart = (double*)f.ws_alloc((size_t)grid * sizeof(double));
  if (!table || !gpart || !epart) return NF_ERR_OOM;
  hw_prep_kernel<T><<<L, 128, 0, f.stream>>>((const T*)theta_dev, f.d_ew_meta, L, d, table);
  NF_LAUNCH_CHECK();
  HwArgs<T> a{};
  a.z0 = (const T*)z0_dev; a.table = table; a.kinds = f.d_ew_kinds;
  a.base = f.base_is_standard ? nullptr : (const T*)f.d_base;
  a.base_c0 = (T)f.base_c0;
  if (tgt) a.tp = tgt->params<T>();
  if (f.score_target) a.sp = f.score_target->params<T>();
  a.y_out = (T*)y_out; a.ld_out = (T*)ld_out; a.terms_out = (T*)terms_out;
  a.gpart = gpart; a.epart = epart; a.N = N; a.L = L; a.d = d;
  a.flags = (tgt ? HW_TARGET : 0) | (want_grad ? HW_GRAD : 0) | (y_out ? HW_WRITE_Y : 0) | (ld_out ? HW_WRITE_LD : 0) |
            (terms_out ? HW_WRITE_TERMS : 0) | (z0_dev ? 0 : HW_GEN_Z0);
  a.seed = seed; a.row0 = f.draw_row_offset;
  f.prof.begin("hmc_warp_flow", f.stream);
  hw_flow_kernel<T><<<grid, HW_THREADS, smem, f.stream>>>(a);
  f.prof.end(f.stream);
  NF_LAUNCH_CHECK();
  if (gsum_dev) {
    hw_finalize_kernel<T><<<L, 128, 0, f.stream>>>((const T*)theta_dev, f.d_ew_meta, L, d, gpart, epart, grid, N, f.P, want_grad ? 1 : 0, 0, gsum_dev);
    NF_LAUNCH_CHECK();
  }
  return NF_OK;
}

template <typename T>
int hmc_warp_inverse(Flow& f, const void* theta_dev, int64_t N, const void* y_dev, bool head, bool want_grad, void* x_out, void* ld_out,
                     void* terms_out, double* gsum_dev) {
  const int L = (int)f.layers.size(), d = f.dim;
  const size_t smem = want_grad ? (size_t)L * 2 * d * sizeof(T) : 16;
  if (smem > 200 * 1024) {
    set_error("Hamiltonian flow with %d layers over dim %d needs %zu B of shared memory for its gradient sums (limit 200 KiB)", L, d, smem);
    return NF_ERR_UNSUPPORTED;
  }
  NF_CUDA(cudaFuncSetAttribute(hw_inv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int max_blocks = 0;
  NF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, hw_inv_kernel<T>, HW_THREADS, smem));
  if (max_blocks < 1) max_blocks = 1;
  const int grid = (int)std::min<int64_t>(ceil_div(N, (int64_t)(HW_THREADS / 32)), (int64_t)kNumSMs * std::min(max_blocks, 4));
  T* table = (T*)f.ws_alloc((size_t)L * (4 + 2 * d) * sizeof(T));
  T* gpart = want_grad ? (T*)f.ws_alloc((size_t)grid * L * 2 * d * sizeof(T)) : nullptr;
  double* epart = (double*)f.ws_alloc((size_t)grid * sizeof(double));
  if (!table || !epart || (want_grad && !gpart)) return NF_ERR_OOM;
  hw_prep_kernel<T><<<L, 128, 0, f.stream>>>((const T*)theta_dev, f.d_ew_meta, L, d, table);
  NF_LAUNCH_CHECK();
  HwArgs<T> a{};
  a.z0 = (const T*)y_dev; a.table = table; a.kinds = f.d_ew_kinds;
  a.base = f.base_is_standard ? nullptr : (const T*)f.d_base;
  a.base_c0 = (T)f.base_c0;
  if (f.score_target) a.sp = f.score_target->params<T>();
  a.y_out = (T*)x_out; a.ld_out = (T*)ld_out; a.terms_out = (T*)terms_out;
  a.gpart = gpart; a.epart = epart; a.N = N; a.L = L; a.d = d;
  a.flags = ((head || want_grad) ? HW_TARGET : 0) | (want_grad ? HW_GRAD : 0) | (x_out ? HW_WRITE_Y : 0) | (ld_out ? HW_WRITE_LD : 0) |
            (terms_out ? HW_WRITE_TERMS : 0);
  f.prof.begin("hmc_warp_inverse", f.stream);
  hw_inv_kernel<T><<<grid, HW_THREADS, smem, f.stream>>>(a);
  f.prof.end(f.stream);
  NF_LAUNCH_CHECK();
  if (gsum_dev) {
    hw_finalize_kernel<T><<<want_grad ? L : 1, 128, 0, f.stream>>>((const T*)theta_dev, f.d_ew_meta, want_grad ? L : 0, d, gpart, epart, grid, N,
                                                                   f.P, want_grad ? 1 : 0, 1, gsum_dev);
    NF_LAUNCH_CHECK();
  }
  return NF_OK;
}

template int hmc_warp_run<float>(Flow&, const Target*, const void*, int64_t, const void*, uint64_t, bool, void*, void*, void*, double*);
template int hmc_warp_run<double>(Flow&, const Target*, const void*, int64_t, const void*, uint64_t, bool, void*, void*, void*, double*);

}  // namespace nf

namespace nf {
template int hmc_warp_inverse<float>(Flow&, const void*, int64_t, const void*, bool, bool, void*, void*, void*, double*);
template int hmc_warp_inverse<double>(Flow&, const void*, int64_t, const void*, bool, bool, void*, void*, void*, double*);
}  // namespace nf
