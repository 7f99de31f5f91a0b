// Device-side target log-densities and scores (d logp / dy) -- SURVEY section 8 row a16, kernel K6.
// Formulas follow reference example/targets/{banana,neal_funnel,warped_gaussian,cross}.jl (App. B).
#pragma once
#include "common.cuh"

namespace nf {

template <typename T> struct TargetParams {
  int kind;        // nf_target_kind
  int dim;
  T p0, p1;        // scalar parameters
  T c0;            // precomputed constant part of logp
  const T* vec;    // DiagNormal: mu[dim] then sigma[dim]; LogReg: X[n][dim] then y[n] (device memory); p0 = sigma0, n_data = n
  int n_data;      // LogReg: number of observations
  int joint;       // 1: logp(z) = logp_inner(z[0:dim]) + sum logN(z[dim:2dim]; 0, 1)  (demo_hamiltonian_flow.jl:117-124)
};

#define NF_LOG2PI 1.8378770664093454835606594728112
#define NF_LOGREG_MAX_DIM 256

// z, g: arrays of length >= d (registers when DP > 0 and loops unroll, else any memory).
// Returns logp(z) and writes g = dlogp/dz.
// LR = false compiles the (data-set walking) logistic-regression case out: the fused elementwise kernels are instantiated both
// ways so that the common targets keep their register budget.
template <typename T, int DP, bool LR = true>
__device__ __forceinline__ T target_logp_score(const TargetParams<T>& tp, const T* z, T* g) {
  using N = Num<T>;
  const int d = tp.dim;
  constexpr int UB = DP > 0 ? DP : 1 << 30;
  switch (tp.kind) {
    case NF_TARGET_BANANA: {  // banana.jl:77-83: b = p0, var = p1
      const T b = tp.p0, v = tp.p1;
      const T u1 = z[0];
      const T u2 = z[1] + b * u1 * u1 - v * b;
      T q = u1 * u1 / v + u2 * u2;
      g[0] = -u1 / v - u2 * (2 * b * u1);
      g[1] = -u2;
#pragma unroll
      for (int k = 2; k < UB; ++k) {
        if (k >= d) break;
        q += z[k] * z[k];
        g[k] = -z[k];
      }
      return tp.c0 - q / 2;
    }
    case NF_TARGET_FUNNEL: {  // neal_funnel.jl:54-72: mu = p0, sigma = p1
      const T mu = tp.p0, sg = tp.p1;
      const T x1 = z[0];
      const T a = N::exp(-x1);
      T ss = 0;
#pragma unroll
      for (int k = 1; k < UB; ++k) {
        if (k >= d) break;
        ss += z[k] * z[k];
        g[k] = -a * z[k];
      }
      g[0] = (mu - x1) / (sg * sg) - T(d - 1) / 2 + a * ss / 2;
      return tp.c0 - (x1 - mu) * (x1 - mu) / (2 * sg * sg) - T(d - 1) / 2 * x1 - a * ss / 2;
    }
    case NF_TARGET_WARPED_GAUSS: {  // warped_gaussian.jl:54-68,81-87: sigma1 = p0, sigma2 = p1
      const T x = z[0], y = z[1];
      const T r2 = x * x + y * y;
      const T r = N::sqrt(r2);
      const T th = N::atan2(y, x) + r / 2;
      T sn, cs;
      N::sincos(th, &sn, &cs);
      const T z1 = r * cs, z2 = r * sn;
      const T i1 = 1 / (tp.p0 * tp.p0), i2 = 1 / (tp.p1 * tp.p1);
      const T lp = tp.c0 - (z1 * z1 * i1 + z2 * z2 * i2) / 2 + N::log(r);
      // d/dz1, d/dz2 of the quadratic, chained through (r, theta)
      const T gz1 = -z1 * i1, gz2 = -z2 * i2;
      const T g_r = gz1 * cs + gz2 * sn;          // dz/dr at fixed theta
      const T g_th = -gz1 * z2 + gz2 * z1;        // dz/dtheta
      const T inv_r = 1 / r, inv_r2 = 1 / r2;
      // r_x = x/r, r_y = y/r ; theta_x = -y/r^2 + x/(2r), theta_y = x/r^2 + y/(2r)
      g[0] = g_r * x * inv_r + g_th * (-y * inv_r2 + x * inv_r / 2) + x * inv_r2;
      g[1] = g_r * y * inv_r + g_th * (x * inv_r2 + y * inv_r / 2) + y * inv_r2;
      return lp;
    }
    case NF_TARGET_CROSS: {  // cross.jl:30-38 per 2-D block; mu = p0, sigma = p1
      const T mu = tp.p0, sg = tp.p1;
      const T isg2 = 1 / (sg * sg);
      T lp = 0;
#pragma unroll
      for (int j = 0; j < UB; j += 2) {
        if (j >= d) break;
        const T a = z[j], b = z[j + 1];
        // components: ([0,mu],(sg,1)) ([-mu,1],(1,sg)) ([mu,1],(1,sg)) ([0,-mu],(sg,1))
        const T e0 = -(a * a * isg2 + (b - mu) * (b - mu)) / 2;
        const T e1 = -((a + mu) * (a + mu) + (b - 1) * (b - 1) * isg2) / 2;
        const T e2 = -((a - mu) * (a - mu) + (b - 1) * (b - 1) * isg2) / 2;
        const T e3 = -(a * a * isg2 + (b + mu) * (b + mu)) / 2;
        const T m = N::max(N::max(e0, e1), N::max(e2, e3));
        const T w0 = N::exp(e0 - m), w1 = N::exp(e1 - m), w2 = N::exp(e2 - m), w3 = N::exp(e3 - m);
        const T s = w0 + w1 + w2 + w3;
        lp += m + N::log(s);
        const T is = 1 / s;
        g[j]     = (w0 * (-a * isg2) + w1 * (-(a + mu)) + w2 * (-(a - mu)) + w3 * (-a * isg2)) * is;
        g[j + 1] = (w0 * (-(b - mu)) + w1 * (-(b - 1) * isg2) + w2 * (-(b - 1) * isg2) + w3 * (-(b + mu))) * is;
      }
      return lp + tp.c0;
    }
    case NF_TARGET_DIAG_NORMAL: {
      T q = 0;
#pragma unroll
      for (int k = 0; k < UB; ++k) {
        if (k >= d) break;
        const T is = 1 / tp.vec[d + k];
        const T u = (z[k] - tp.vec[k]) * is;
        q += u * u;
        g[k] = -u * is;
      }
      return tp.c0 - q / 2;
    }
    case NF_TARGET_LOGREG: if constexpr (LR) {  // u_i = x_i . z ;  logp = sum_i [y_i u_i - softplus(u_i)] - |z|^2 / (2 sigma0^2) + c0
      const T is2 = 1 / (tp.p0 * tp.p0);
      const T* X = tp.vec;
      const T* y = tp.vec + (size_t)tp.n_data * d;
      T lp = 0, q = 0;
      T zz[DP > 0 ? DP : 1];   // the inputs are read before g is written (z and g may alias)
      if (DP > 0) {
#pragma unroll
        for (int k = 0; k < UB; ++k) { if (k >= d) break; zz[k] = z[k]; q += z[k] * z[k]; g[k] = -z[k] * is2; }
        for (int i = 0; i < tp.n_data; ++i) {
          const T* xi = X + (size_t)i * d;
          T u = 0;
#pragma unroll
          for (int k = 0; k < UB; ++k) { if (k >= d) break; u += xi[k] * zz[k]; }
          const T yi = y[i];
          lp += yi * u - softplus_stable<T>(u);
          const T r = yi - sigmoid_stable<T>(u);
#pragma unroll
          for (int k = 0; k < UB; ++k) { if (k >= d) break; g[k] += r * xi[k]; }
        }
      } else {
        // generic-memory variant (heads of the layered path, where z and g may be the SAME row): private copies of the input
        // and of the score accumulator live in (L1-resident) local memory, so the data set is walked once
        T zc[NF_LOGREG_MAX_DIM], acc[NF_LOGREG_MAX_DIM];
        for (int k = 0; k < d; ++k) { zc[k] = z[k]; q += zc[k] * zc[k]; acc[k] = -zc[k] * is2; }
        for (int i = 0; i < tp.n_data; ++i) {
          const T* xi = X + (size_t)i * d;
          T u = 0;
          for (int k = 0; k < d; ++k) u += xi[k] * zc[k];
          const T yi = y[i];
          lp += yi * u - softplus_stable<T>(u);
          const T r = yi - sigmoid_stable<T>(u);
          for (int k = 0; k < d; ++k) acc[k] += r * xi[k];
        }
        for (int k = 0; k < d; ++k) g[k] = acc[k];
      }
      return tp.c0 + lp - q * is2 / 2;
    }
    break;
  }
  return 0;
}

// out = (Hessian of logp at x) * w -- the second-order term the reverse sweep through a LeapFrog layer needs
// (reference example/demo_hamiltonian_flow.jl:49-61 differentiates through `∇logp`).  Supported for the targets
// whose score is smooth and closed-form: Banana, Funnel, DiagNormal.
template <typename T, int DP, bool LR = true>
__device__ __forceinline__ void target_hvp(const TargetParams<T>& tp, const T* x, const T* w, T* out) {
  using N = Num<T>;
  const int d = tp.dim;
  constexpr int UB = DP > 0 ? DP : 1 << 30;
  switch (tp.kind) {
    case NF_TARGET_BANANA: {
      const T b = tp.p0, v = tp.p1;
      const T u2 = x[1] + b * x[0] * x[0] - v * b;
      const T h11 = -1 / v - 2 * b * u2 - 4 * b * b * x[0] * x[0];
      const T h12 = -2 * b * x[0];
      out[0] = h11 * w[0] + h12 * w[1];
      out[1] = h12 * w[0] - w[1];
#pragma unroll
      for (int k = 2; k < UB; ++k) {
        if (k >= d) break;
        out[k] = -w[k];
      }
      return;
    }
    case NF_TARGET_FUNNEL: {
      const T sg = tp.p1;
      const T a = N::exp(-x[0]);
      T ss = 0, xw = 0;
#pragma unroll
      for (int k = 1; k < UB; ++k) {
        if (k >= d) break;
        ss += x[k] * x[k];
        xw += x[k] * w[k];
        out[k] = a * (x[k] * w[0] - w[k]);
      }
      out[0] = (-1 / (sg * sg) - a * ss / 2) * w[0] + a * xw;
      return;
    }
    case NF_TARGET_DIAG_NORMAL: {
#pragma unroll
      for (int k = 0; k < UB; ++k) {
        if (k >= d) break;
        const T is = 1 / tp.vec[d + k];
        out[k] = -w[k] * is * is;
      }
      return;
    }
    case NF_TARGET_LOGREG: if constexpr (LR) {  // H w = -X^T diag(s (1 - s)) X w - w / sigma0^2,  s = sigmoid(X x)
      const T is2 = 1 / (tp.p0 * tp.p0);
      const T* X = tp.vec;
#pragma unroll
      for (int k = 0; k < UB; ++k) { if (k >= d) break; out[k] = -w[k] * is2; }
      for (int i = 0; i < tp.n_data; ++i) {
        const T* xi = X + (size_t)i * d;
        T u = 0, xw = 0;
#pragma unroll
        for (int k = 0; k < UB; ++k) { if (k >= d) break; u += xi[k] * x[k]; xw += xi[k] * w[k]; }
        const T sg = sigmoid_stable<T>(u);
        const T cfac = -sg * (1 - sg) * xw;
#pragma unroll
        for (int k = 0; k < UB; ++k) { if (k >= d) break; out[k] += cfac * xi[k]; }
      }
      return;
    }
  }
}

__host__ __device__ inline bool target_has_hvp(int kind) {
  return kind == NF_TARGET_BANANA || kind == NF_TARGET_FUNNEL || kind == NF_TARGET_DIAG_NORMAL || kind == NF_TARGET_LOGREG;
}

}  // namespace nf
