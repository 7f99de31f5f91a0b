// K1/K2 -- fused forward + logdet + target + backward kernel for flows made only of elementwise
// layers (PlanarLayer, RadialLayer, Shift, Scale): SURVEY section 8 rows a9, a10, a15, a16.
//
// One launch turns Z0 [N x d] into the ELBO partial sums and the per-layer gradient partial sums;
// nothing of size N is written.  Each thread owns S samples whose state lives in registers, the
// per-layer parameter table is staged in shared memory, the per-sample/per-layer scalar needed by
// the backward sweep (tanh(a) for planar, r for radial) is stashed in shared memory, and the
// per-parameter gradient sums are reduced warp-shuffle -> shared atomics -> one partial row per CTA
// (finished deterministically by ew_finalize_kernel).
//
// Math (reference-side definitions):
//   planar  (Bijectors.PlanarLayer, restated in reference test/ext/CUDA/cuda.jl:12-30; App. A.1)
//   radial  (Bijectors.RadialLayer; App. A.2)
//   elbo_j = logp(T(x_j)) - log q0(x_j) + logdet_j          (reference src/objectives/elbo.jl:4-7,65-70)
#pragma once
#include "flow.hpp"
#include "targets.cuh"
#include "base_dense.cuh"

namespace nf {

// per-layer table entry: [c0 c1 c2 c3 | v0[DP] | v1[DP]]
//   planar: c0 = b, c1 = m = w.u_hat ; v0 = w, v1 = u_hat
//   radial: c0 = alpha, c1 = beta_hat ; v0 = z0
//   shift : v0 = a
//   scale : c0 = sum log|a| ; v0 = a, v1 = 1/a
//   momentum affine (z = [x, rho], h = DP/2): c0 = sum log|a| ; v0 = [0.., b], v1 = [1.., a]
//   leapfrog: c0 = number of steps ; v0[0:h] = eps = exp(log_eps)
template <int DP> constexpr int ew_stride() { return 4 + 2 * DP; }
template <int DP> constexpr int ew_nacc() { return 2 * DP + 2; }

template <typename T, int DP>
__device__ __forceinline__ void ew_prep_body(const T* __restrict__ theta, const EwLayerMeta* __restrict__ meta, int l, int d,
                                             T* __restrict__ table) {
  using N = Num<T>;
  const T* p = theta + meta[l].theta_off;
  T* e = table + (size_t)l * ew_stride<DP>();
  for (int k = 0; k < ew_stride<DP>(); ++k) e[k] = 0;
  T* v0 = e + 4;
  T* v1 = e + 4 + DP;
  switch (meta[l].kind) {
    case NF_PLANAR: {  // theta: w(d), u(d), b
      const T* w = p; const T* u = p + d;
      T s = 0, n = 0;
      for (int k = 0; k < d; ++k) { s += w[k] * u[k]; n += w[k] * w[k]; }
      const T kappa = (softplus_stable<T>(-s) - 1) / n;
      for (int k = 0; k < d; ++k) { v0[k] = w[k]; v1[k] = u[k] + kappa * w[k]; }
      e[0] = p[2 * d];
      e[1] = softplus_stable<T>(s) - 1;
      break;
    }
    case NF_RADIAL: {  // theta: alpha_, beta, z0(d)
      const T alpha = softplus_stable<T>(p[0]);
      e[0] = alpha;
      e[1] = -alpha + softplus_stable<T>(p[1]);
      for (int k = 0; k < d; ++k) v0[k] = p[2 + k];
      break;
    }
    case NF_SHIFT:
      for (int k = 0; k < d; ++k) v0[k] = p[k];
      break;
    case NF_SCALE: {
      T sl = 0;
      for (int k = 0; k < d; ++k) { v0[k] = p[k]; v1[k] = 1 / p[k]; sl += N::log(N::abs(p[k])); }
      for (int k = d; k < DP; ++k) { v0[k] = 1; v1[k] = 1; }
      e[0] = sl;
      break;
    }
    case NF_MOMENTUM_AFFINE: {  // theta: b(h), a(h); d == DP == 2h
      const int h = d / 2;
      T sl = 0;
      for (int k = 0; k < DP; ++k) { v0[k] = 0; v1[k] = 1; }
      for (int k = 0; k < h; ++k) { v0[h + k] = p[k]; v1[h + k] = p[h + k]; sl += N::log(N::abs(p[h + k])); }
      e[0] = sl;
      break;
    }
    case NF_LEAPFROG: {  // theta: log_eps(h)
      const int h = d / 2;
      for (int k = 0; k < h; ++k) v0[k] = N::exp(p[k]);
      e[0] = (T)meta[l].aux;
      break;
    }
  }
}

template <typename T, int DP>
__global__ void ew_prep_kernel(const T* __restrict__ theta, const EwLayerMeta* __restrict__ meta, int L, int d,
                               T* __restrict__ table) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < L) ew_prep_body<T, DP>(theta, meta, l, d, table);
}

// ---------------------------------------------------------------------------------------------
// LeapFrog bijector of the Hamiltonian flow (reference example/demo_hamiltonian_flow.jl:27-91):
//   rho += eps/2 .* s(x);  { x += eps .* rho;  rho += eps .* s(x) } x (L-1);  x += eps .* rho;  rho += eps/2 .* s(x)
// with s = score of the target.  Volume preserving (logdet 0, :84-91); the inverse is the same map with -eps (:63-82).
// The reverse sweep needs no stash: each elementary update is undone exactly (up to rounding) while its adjoint is
// applied, the second-order term being a Hessian-vector product of the target.
// ---------------------------------------------------------------------------------------------
template <typename T, int HD, bool LR>
__device__ __forceinline__ void leapfrog_apply(const TargetParams<T>& sp, T* x, T* v, const T* eps, T sgn, int nsteps) {
  T g[HD];
#pragma unroll
  for (int k = 0; k < HD; ++k) g[k] = 0;
  target_logp_score<T, HD, LR>(sp, x, g);
#pragma unroll
  for (int k = 0; k < HD; ++k) v[k] += sgn * eps[k] / 2 * g[k];
  for (int it = 0; it < nsteps; ++it) {
#pragma unroll
    for (int k = 0; k < HD; ++k) x[k] += sgn * eps[k] * v[k];
    target_logp_score<T, HD, LR>(sp, x, g);
    const T c = (it == nsteps - 1) ? sgn / 2 : sgn;
#pragma unroll
    for (int k = 0; k < HD; ++k) v[k] += c * eps[k] * g[k];
  }
}

// (x, v): OUTPUT state of leapfrog_apply(sgn) on entry, its input state on exit; (gx, gv): adjoints of the output on
// entry, of the input on exit; geps += d/d(eps) (the derivative w.r.t. the signed step is folded in through sgn).
template <typename T, int HD, bool LR>
__device__ __forceinline__ void leapfrog_backward(const TargetParams<T>& sp, T* x, T* v, T* gx, T* gv, const T* eps,
                                                  T sgn, int nsteps, T* geps) {
  T g[HD], w[HD], hw[HD];
#pragma unroll
  for (int k = 0; k < HD; ++k) { g[k] = 0; hw[k] = 0; }
  for (int it = nsteps - 1; it >= 0; --it) {
    // undo  v += c eps s(x)
    const T c = (it == nsteps - 1) ? sgn / 2 : sgn;
    target_logp_score<T, HD, LR>(sp, x, g);
#pragma unroll
    for (int k = 0; k < HD; ++k) {
      v[k] -= c * eps[k] * g[k];
      geps[k] += c * gv[k] * g[k];
      w[k] = c * eps[k] * gv[k];
    }
    target_hvp<T, HD, LR>(sp, x, w, hw);
#pragma unroll
    for (int k = 0; k < HD; ++k) gx[k] += hw[k];
    // undo  x += eps v
#pragma unroll
    for (int k = 0; k < HD; ++k) {
      x[k] -= sgn * eps[k] * v[k];
      geps[k] += sgn * gx[k] * v[k];
      gv[k] += sgn * eps[k] * gx[k];
    }
  }
  // undo the opening half kick
  target_logp_score<T, HD, LR>(sp, x, g);
#pragma unroll
  for (int k = 0; k < HD; ++k) {
    v[k] -= sgn * eps[k] / 2 * g[k];
    geps[k] += sgn / 2 * gv[k] * g[k];
    w[k] = sgn * eps[k] / 2 * gv[k];
  }
  target_hvp<T, HD, LR>(sp, x, w, hw);
#pragma unroll
  for (int k = 0; k < HD; ++k) gx[k] += hw[k];
}

// EW_EXT_GRAD: a run of elementwise layers INSIDE a layered (coupling) flow -- no target; the backward sweep starts from the
// caller's d/dy (g_in) and per-sample d/dlogdet (gld, nullptr = 1) and leaves d/dx in g_out.  EW_ADD_LD: ld_out[j] += logdet.
enum : int { EW_GRAD = 1, EW_TARGET = 2, EW_WRITE_Y = 4, EW_WRITE_LD = 8, EW_WRITE_TERMS = 16, EW_GEN_Z0 = 32, EW_EXT_GRAD = 64, EW_ADD_LD = 128 };

template <typename T> struct EwArgs {
  const T* z0;          // [N, d] (ignored with EW_GEN_Z0)
  const T* table;       // [L, stride]
  const int* kinds;     // [L]
  const T* base;        // mu[d], sigma[d] or nullptr
  T base_c0;            // -d/2 log2pi - sum log sigma
  TargetParams<T> tp;
  TargetParams<T> sp;   // score target of the LeapFrog layers (dim d/2)
  T* y_out;             // [N, d]
  T* ld_out;            // [N]
  T* terms_out;         // [N]
  T* gpart;             // [grid, L * nacc]
  double* epart;        // [grid]
  int64_t N;
  int L, d, flags;
  uint64_t seed;
  int64_t row0;         // global row of sample 0 (Philox draws of a data-parallel shard)
  const T* lq0;         // optional per-sample log q0(x0) (full-covariance base: computed by base_dense_kernel; base == nullptr then)
  const T* g_in;        // EW_EXT_GRAD: [N, d] d/dy
  T* g_out;             // EW_EXT_GRAD: [N, d] d/dx (may alias g_in)
  const T* gld;         // EW_EXT_GRAD: [N] d/dlogdet or nullptr (= 1)
};

template <typename T, int DP, int S, bool LR>
__device__ __forceinline__ void ew_flow_body(const EwArgs<T>& a, const int bid, const int nblk) {
  using N_ = Num<T>;
  constexpr int STR = ew_stride<DP>();
  constexpr int NACC = ew_nacc<DP>();
  constexpr int HD = DP / 2;
  const int L = a.L, d = a.d, tid = threadIdx.x, nthr = blockDim.x;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* s_tab = reinterpret_cast<T*>(smem_raw);                 // L*STR
  const int nwarp = nthr >> 5;
  T* s_acc = s_tab + (size_t)L * STR;                        // nwarp*L*NACC: one row of sums per warp (lane 0 adds, no atomics)
  T* s_stash = s_acc + (size_t)nwarp * L * NACC;             // L*S*nthr
  int* s_kind = reinterpret_cast<int*>(s_stash + (size_t)L * S * nthr);  // L
  for (int i = tid; i < L * STR; i += nthr) s_tab[i] = a.table[i];
  for (int i = tid; i < nwarp * L * NACC; i += nthr) s_acc[i] = 0;
  for (int i = tid; i < L; i += nthr) s_kind[i] = a.kinds[i];
  __syncthreads();

  const bool want_grad = a.flags & EW_GRAD;
  T elbo_local = 0;
  const int64_t group = (int64_t)nthr * S;
  const int64_t ngroups = (a.N + group - 1) / group;

  for (int64_t gi = bid; gi < ngroups; gi += nblk) {
    T z[S][DP], ld[S], lq[S];
    bool live[S];
    T wl[S];              // d(objective)/d(logdet) of the sample: 1 for the ELBO, the caller's weight in EW_EXT_GRAD mode, 0 if not live
    // ---- load base draws, base log-density (a15) ----
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int64_t j = gi * group + (int64_t)s * nthr + tid;
      live[s] = j < a.N;
      const int64_t jj = live[s] ? j : 0;
      T q = 0;
      if (!(a.flags & EW_GEN_Z0) && DP == 2 && d == 2 && sizeof(T) == 4) {
        const float2 v = reinterpret_cast<const float2*>(a.z0)[jj];
        z[s][0] = v.x; z[s][1] = v.y;
      } else {
#pragma unroll
        for (int k = 0; k < DP; ++k) {
          if (k < d) z[s][k] = (a.flags & EW_GEN_Z0) ? philox_randn<T>(a.seed, (uint64_t)(a.row0 + jj) * d + k) : a.z0[jj * d + k];
          else z[s][k] = 0;
        }
      }
      if (a.base) {
#pragma unroll
        for (int k = 0; k < DP; ++k)
          if (k < d) {
            if (a.flags & EW_GEN_Z0) { q += z[s][k] * z[s][k]; z[s][k] = z[s][k] * a.base[d + k] + a.base[k]; }
            else { const T u = (z[s][k] - a.base[k]) / a.base[d + k]; q += u * u; }
          }
      } else {
#pragma unroll
        for (int k = 0; k < DP; ++k) q += z[s][k] * z[s][k];
      }
      lq[s] = a.lq0 ? a.lq0[jj] : a.base_c0 - q / 2;
      ld[s] = 0;
      wl[s] = live[s] ? (((a.flags & EW_EXT_GRAD) && a.gld) ? a.gld[jj] : T(1)) : T(0);
    }
    {   // a warp without a single live sample (ragged last group, tiny batches) has nothing to add to any sum
      bool any_live = false;
#pragma unroll
      for (int s = 0; s < S; ++s) any_live |= live[s];
      if (!__any_sync(0xffffffffu, any_live)) continue;
    }
    // ---- forward sweep: layers applied last-to-first (create_flow, reference src/flows/utils.jl:23-26) ----
    for (int l = L - 1; l >= 0; --l) {
      const T* e = s_tab + (size_t)l * STR;
      const int kind = s_kind[l];
      if (kind == NF_PLANAR) {
        const T b = e[0], m = e[1];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          T dot = b;
#pragma unroll
          for (int k = 0; k < DP; ++k) dot += e[4 + k] * z[s][k];
          const T t = N_::tanh(dot);
#pragma unroll
          for (int k = 0; k < DP; ++k) z[s][k] += e[4 + DP + k] * t;
          ld[s] += N_::log1p(m * (1 - t * t));
          s_stash[((size_t)l * S + s) * nthr + tid] = t;
        }
      } else if (kind == NF_RADIAL) {
        const T alpha = e[0], bh = e[1];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          T r2 = 0, df[DP];
#pragma unroll
          for (int k = 0; k < DP; ++k) { df[k] = (k < d) ? z[s][k] - e[4 + k] : T(0); r2 += df[k] * df[k]; }
          const T r = N_::sqrt(r2);
          const T h = 1 / (alpha + r);
          const T g = bh * h;
#pragma unroll
          for (int k = 0; k < DP; ++k) z[s][k] += g * df[k];
          ld[s] += T(d - 1) * N_::log1p(g) + N_::log1p(g - g * h * r);
          s_stash[((size_t)l * S + s) * nthr + tid] = r;
        }
      } else if (kind == NF_SHIFT) {
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int k = 0; k < DP; ++k) z[s][k] += e[4 + k];
      } else if (kind == NF_SCALE) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll
          for (int k = 0; k < DP; ++k) z[s][k] *= e[4 + k];
          ld[s] += e[0];
        }
      } else if (kind == NF_MOMENTUM_AFFINE) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll
          for (int k = HD; k < DP; ++k) z[s][k] = z[s][k] * e[4 + DP + k] + e[4 + k];
          ld[s] += e[0];
        }
      } else {  // NF_LEAPFROG
        T eps[HD];
#pragma unroll
        for (int k = 0; k < HD; ++k) eps[k] = e[4 + k];
        const int nst = (int)e[0];
#pragma unroll
        for (int s = 0; s < S; ++s) leapfrog_apply<T, HD, LR>(a.sp, z[s], z[s] + HD, eps, T(1), nst);
      }
    }
    // ---- outputs of a pure forward pass ----
    if (a.flags & (EW_WRITE_Y | EW_WRITE_LD)) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int64_t j = gi * group + (int64_t)s * nthr + tid;
        if (!live[s]) continue;
        if (a.flags & EW_WRITE_Y)
#pragma unroll
          for (int k = 0; k < DP; ++k)
            if (k < d) a.y_out[j * d + k] = z[s][k];
        if (a.flags & EW_WRITE_LD) a.ld_out[j] = ld[s];
      }
    }
    if (a.flags & EW_ADD_LD) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int64_t j = gi * group + (int64_t)s * nthr + tid;
        if (live[s]) a.ld_out[j] += ld[s];
      }
    }
    if (!(a.flags & (EW_TARGET | EW_EXT_GRAD))) continue;
    // ---- target log-density + score (a16), ELBO term ----
    T gy[S][DP];
    if (a.flags & EW_EXT_GRAD) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int64_t j = gi * group + (int64_t)s * nthr + tid;
#pragma unroll
        for (int k = 0; k < DP; ++k) gy[s][k] = (live[s] && k < d) ? a.g_in[j * d + k] : T(0);
      }
    } else
#pragma unroll
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int k = 0; k < DP; ++k) gy[s][k] = 0;
      T lp;
      if (a.tp.joint) {   // logp(x) + sum logN(rho; 0, 1)   (demo_hamiltonian_flow.jl:117-124); d == DP here
        lp = target_logp_score<T, HD, LR>(a.tp, z[s], gy[s]);
        T q = 0;
#pragma unroll
        for (int k = HD; k < DP; ++k) { q += z[s][k] * z[s][k]; gy[s][k] = -z[s][k]; }
        lp -= q / 2 + T(HD) * T(NF_LOG2PI / 2);
      } else {
        lp = target_logp_score<T, DP, LR>(a.tp, z[s], gy[s]);
      }
      const T term = lp - lq[s] + ld[s];
      if (live[s]) {
        elbo_local += term;
        if (a.flags & EW_WRITE_TERMS) a.terms_out[gi * group + (int64_t)s * nthr + tid] = term;
      } else {
#pragma unroll
        for (int k = 0; k < DP; ++k) gy[s][k] = 0;
      }
    }
    if (!want_grad) continue;
    // ---- backward sweep (reverse of application order); dELBO/dlogdet = 1 per live sample ----
    const int lane = tid & 31;
    for (int l = 0; l < L; ++l) {
      const T* e = s_tab + (size_t)l * STR;
      T* acc = s_acc + ((size_t)(tid >> 5) * L + l) * NACC;
      const int kind = s_kind[l];
      if (kind == NF_PLANAR) {
        const T m = e[1];
        T ga[S], tt[S];
        T g_m = 0, g_b = 0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const T t = s_stash[((size_t)l * S + s) * nthr + tid];
          const T w8 = wl[s];
          const T psi = 1 - t * t, den = 1 + m * psi;
          T ug = 0;
#pragma unroll
          for (int k = 0; k < DP; ++k) { z[s][k] -= e[4 + DP + k] * t; ug += e[4 + DP + k] * gy[s][k]; }
          const T gt = ug - w8 * 2 * t * m / den;
          ga[s] = psi * gt; tt[s] = t;
          g_m += w8 * psi / den; g_b += ga[s];
        }
#pragma unroll
        for (int k = 0; k < DP; ++k) {
          T gu = 0, gw = 0;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            gu += tt[s] * gy[s][k];
            gw += z[s][k] * ga[s];
            gy[s][k] += e[4 + k] * ga[s];
          }
          if (k < d) {
            gu = warp_sum(gu); gw = warp_sum(gw);
            if (lane == 0) { acc[k] += gu; acc[DP + k] += gw; }
          }
        }
        g_m = warp_sum(g_m); g_b = warp_sum(g_b);
        if (lane == 0) { acc[2 * DP] += g_m; acc[2 * DP + 1] += g_b; }
      } else if (kind == NF_RADIAL) {
        const T alpha = e[0], bh = e[1];
        T g_al = 0, g_bh = 0, gz0[DP];
#pragma unroll
        for (int k = 0; k < DP; ++k) gz0[k] = 0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const T r = s_stash[((size_t)l * S + s) * nthr + tid];
          const T w8 = wl[s];
          const T h = 1 / (alpha + r), g = bh * h;
          const T ifac = 1 / (1 + g);
          T df[DP], gd = 0;
#pragma unroll
          for (int k = 0; k < DP; ++k) {
            df[k] = (k < d) ? (z[s][k] - e[4 + k]) * ifac : T(0);
            gd += gy[s][k] * df[k];
            z[s][k] = (k < d) ? e[4 + k] + df[k] : T(0);
          }
          const T A = 1 + g, Bq = 1 + alpha * bh * h * h;
          const T dld_dh = T(d - 1) * bh / A + 2 * alpha * bh * h / Bq;
          // through r: g(r) = bh*h(r), h' = -h^2
          const T coef_r = (gd * (-bh * h * h) + w8 * (-h * h) * dld_dh);
          g_al += gd * (-bh * h * h) + w8 * ((-h * h) * dld_dh + bh * h * h / Bq);
          g_bh += gd * h + w8 * (T(d - 1) * h / A + alpha * h * h / Bq);
          const T ir = r > 0 ? 1 / r : T(0);
#pragma unroll
          for (int k = 0; k < DP; ++k) {
            const T gnew = (1 + g) * gy[s][k] + coef_r * df[k] * ir;
            gz0[k] += gy[s][k] - gnew;
            gy[s][k] = gnew;
          }
        }
#pragma unroll
        for (int k = 0; k < DP; ++k)
          if (k < d) {
            const T v = warp_sum(gz0[k]);
            if (lane == 0) acc[k] += v;
          }
        g_al = warp_sum(g_al); g_bh = warp_sum(g_bh);
        if (lane == 0) { acc[2 * DP] += g_al; acc[2 * DP + 1] += g_bh; }
      } else if (kind == NF_SHIFT) {
#pragma unroll
        for (int k = 0; k < DP; ++k) {
          T v = 0;
#pragma unroll
          for (int s = 0; s < S; ++s) { v += gy[s][k]; z[s][k] -= e[4 + k]; }
          if (k < d) { v = warp_sum(v); if (lane == 0) acc[k] += v; }
        }
      } else if (kind == NF_SCALE) {
#pragma unroll
        for (int k = 0; k < DP; ++k) {
          T v = 0;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            z[s][k] *= e[4 + DP + k];
            v += gy[s][k] * z[s][k];
            gy[s][k] *= e[4 + k];
          }
          if (k < d) { v = warp_sum(v); if (lane == 0) acc[k] += v; }
        }
      } else if (kind == NF_MOMENTUM_AFFINE) {
#pragma unroll
        for (int k = HD; k < DP; ++k) {
          T gb = 0, ga = 0;
          const T ia = 1 / e[4 + DP + k];
#pragma unroll
          for (int s = 0; s < S; ++s) {
            z[s][k] = (z[s][k] - e[4 + k]) * ia;
            gb += gy[s][k];
            ga += gy[s][k] * z[s][k];
            gy[s][k] *= e[4 + DP + k];
          }
          gb = warp_sum(gb); ga = warp_sum(ga);
          if (lane == 0) { acc[k] += gb; acc[DP + k] += ga; }
        }
      } else {  // NF_LEAPFROG
        T eps[HD], ge[HD];
#pragma unroll
        for (int k = 0; k < HD; ++k) { eps[k] = e[4 + k]; ge[k] = 0; }
        const int nst = (int)e[0];
#pragma unroll
        for (int s = 0; s < S; ++s)
          leapfrog_backward<T, HD, LR>(a.sp, z[s], z[s] + HD, gy[s], gy[s] + HD, eps, T(1), nst, ge);
#pragma unroll
        for (int k = 0; k < HD; ++k) {
          const T v = warp_sum(ge[k]);
          if (lane == 0) acc[k] += v;
        }
      }
    }
    if (a.flags & EW_EXT_GRAD) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int64_t j = gi * group + (int64_t)s * nthr + tid;
        if (!live[s]) continue;
#pragma unroll
        for (int k = 0; k < DP; ++k)
          if (k < d) a.g_out[j * d + k] = gy[s][k];
      }
    }
  }
  // ---- per-CTA partials ----
  __syncthreads();
  if (a.gpart)
    for (int i = tid; i < L * NACC; i += nthr) {
      T t = 0;
      for (int w = 0; w < nwarp; ++w) t += s_acc[(size_t)w * L * NACC + i];
      a.gpart[(size_t)bid * L * NACC + i] = t;
    }
  if (a.epart) {
    double ev = warp_sum((double)elbo_local);
    __shared__ double s_e[32];
    if ((tid & 31) == 0) s_e[tid >> 5] = ev;
    __syncthreads();
    if (tid == 0) {
      double t = 0;
      for (int w = 0; w < (nthr + 31) / 32; ++w) t += s_e[w];
      a.epart[bid] = t;
    }
  }
}

template <typename T, int DP, int S, bool LR>
__global__ void __launch_bounds__(128) ew_flow_kernel(EwArgs<T> a) {
  ew_flow_body<T, DP, S, LR>(a, (int)blockIdx.x, (int)gridDim.x);
}


// ---------------------------------------------------------------------------------------------
// Inverse direction (SURVEY row a5 / kernel K5): x0 = T^{-1}(y), logdet of the inverse, and for the
// forward-KL objective `loglikelihood` (reference src/objectives/loglikelihood.jl:26-33, Bijectors
// logpdf(td, y) = logpdf(q0, x0) + logdet_inv) the gradient w.r.t. theta by implicit differentiation:
//   z = f^{-1}(y; th):  dL/dy = J^{-T} g,  dL/dth = -(df/dth)^T J^{-T} g - dlogdet_fwd/dth,
//   g = dL/dz - grad_z logdet_fwd(z),  J = df/dz  (rank-one updates of the identity -> Sherman-Morrison).
// Planar inverse: scalar root find of  a + m tanh(a + b) = w.y  (monotone since m = w.u_hat > -1; App. A.1),
// safeguarded Newton inside the bracket [w.y - |m|, w.y + |m|].  Radial inverse: closed form (App. A.2).
// No stash is needed: the backward sweep walks the FORWARD maps from x0 back to y.
// ---------------------------------------------------------------------------------------------
template <typename T, int DP, int S, bool LR>
__global__ void __launch_bounds__(128) ew_inv_flow_kernel(EwArgs<T> a) {
  using N_ = Num<T>;
  constexpr int STR = ew_stride<DP>();
  constexpr int NACC = ew_nacc<DP>();
  constexpr int HD = DP / 2;
  const int L = a.L, d = a.d, tid = threadIdx.x, nthr = blockDim.x;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* s_tab = reinterpret_cast<T*>(smem_raw);
  const int nwarp = nthr >> 5;
  T* s_acc = s_tab + (size_t)L * STR;                        // nwarp*L*NACC (see ew_flow_body)
  int* s_kind = reinterpret_cast<int*>(s_acc + (size_t)nwarp * L * NACC);
  for (int i = tid; i < L * STR; i += nthr) s_tab[i] = a.table[i];
  for (int i = tid; i < nwarp * L * NACC; i += nthr) s_acc[i] = 0;
  for (int i = tid; i < L; i += nthr) s_kind[i] = a.kinds[i];
  __syncthreads();
  const bool want_grad = a.flags & EW_GRAD;
  const T tol = sizeof(T) == 4 ? T(1e-7) : T(1e-15);
  T obj_local = 0;
  const int lane = tid & 31;
  const int64_t group = (int64_t)nthr * S;
  const int64_t ngroups = (a.N + group - 1) / group;
  for (int64_t gi = blockIdx.x; gi < ngroups; gi += gridDim.x) {
    T z[S][DP], ld[S], wl[S];
    bool live[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int64_t j = gi * group + (int64_t)s * nthr + tid;
      live[s] = j < a.N;
      const int64_t jj = live[s] ? j : 0;
#pragma unroll
      for (int k = 0; k < DP; ++k) z[s][k] = (k < d) ? a.z0[jj * d + k] : T(0);
      ld[s] = 0;
      wl[s] = live[s] ? (((a.flags & EW_EXT_GRAD) && a.gld) ? a.gld[jj] : T(1)) : T(0);
    }
    // ---- inverse sweep: inverse(f1∘...∘fL) applies f1^{-1} first (theta order) ----
    for (int l = 0; l < L; ++l) {
      const T* e = s_tab + (size_t)l * STR;
      const int kind = s_kind[l];
      if (kind == NF_PLANAR) {
        const T b = e[0], m = e[1];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          T c = 0;
#pragma unroll
          for (int k = 0; k < DP; ++k) c += e[4 + k] * z[s][k];
          const T am = N_::abs(m);
          T lo = c - am, hi = c + am, al = c, t = 0;
          for (int it = 0; it < 60; ++it) {
            t = N_::tanh(al + b);
            const T f = al + m * t - c;
            if (N_::abs(f) <= tol * (1 + N_::abs(c))) break;
            if (f > 0) hi = al; else lo = al;
            T an = al - f / (1 + m * (1 - t * t));
            if (!(an > lo && an < hi)) an = (lo + hi) / 2;
            al = an;
          }
          t = N_::tanh(al + b);
#pragma unroll
          for (int k = 0; k < DP; ++k) z[s][k] -= e[4 + DP + k] * t;
          ld[s] -= N_::log1p(m * (1 - t * t));
        }
      } else if (kind == NF_RADIAL) {
        const T alpha = e[0], bh = e[1];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          T rho2 = 0, df[DP];
#pragma unroll
          for (int k = 0; k < DP; ++k) { df[k] = (k < d) ? z[s][k] - e[4 + k] : T(0); rho2 += df[k] * df[k]; }
          const T rho = N_::sqrt(rho2);
          const T q = (alpha + bh) - rho;
          const T r = (N_::sqrt(q * q + 4 * alpha * rho) - q) / 2;
          const T fac = (alpha + r) / (alpha + bh + r);
#pragma unroll
          for (int k = 0; k < DP; ++k) z[s][k] = (k < d) ? e[4 + k] + fac * df[k] : T(0);
          const T h = 1 / (alpha + r), g = bh * h;
          ld[s] -= T(d - 1) * N_::log1p(g) + N_::log1p(g - g * h * r);
        }
      } else if (kind == NF_SHIFT) {
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int k = 0; k < DP; ++k) z[s][k] -= e[4 + k];
      } else if (kind == NF_SCALE) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll
          for (int k = 0; k < DP; ++k) z[s][k] *= e[4 + DP + k];
          ld[s] -= e[0];
        }
      } else if (kind == NF_MOMENTUM_AFFINE) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll
          for (int k = HD; k < DP; ++k) z[s][k] = (z[s][k] - e[4 + k]) / e[4 + DP + k];
          ld[s] -= e[0];
        }
      } else {  // NF_LEAPFROG: inverse = same map with -eps (demo_hamiltonian_flow.jl:63-82)
        T eps[HD];
#pragma unroll
        for (int k = 0; k < HD; ++k) eps[k] = e[4 + k];
        const int nst = (int)e[0];
#pragma unroll
        for (int s = 0; s < S; ++s) leapfrog_apply<T, HD, LR>(a.sp, z[s], z[s] + HD, eps, T(-1), nst);
      }
    }
    if (a.flags & (EW_WRITE_Y | EW_WRITE_LD)) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int64_t j = gi * group + (int64_t)s * nthr + tid;
        if (!live[s]) continue;
        if (a.flags & EW_WRITE_Y)
#pragma unroll
          for (int k = 0; k < DP; ++k)
            if (k < d) a.y_out[j * d + k] = z[s][k];
        if (a.flags & EW_WRITE_LD) a.ld_out[j] = ld[s];
      }
    }
    if (a.flags & EW_ADD_LD) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int64_t j = gi * group + (int64_t)s * nthr + tid;
        if (live[s]) a.ld_out[j] += ld[s];
      }
    }
    if (!(a.flags & (EW_TARGET | EW_EXT_GRAD))) continue;     // EW_TARGET here means "log-likelihood head"
    // ---- head: logpdf(q0, x0) + logdet_inv ;  g = d logq0 / d x0  (EW_EXT_GRAD: the caller's d/dx instead) ----
    T gz[S][DP];
    if (a.flags & EW_EXT_GRAD) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int64_t j = gi * group + (int64_t)s * nthr + tid;
#pragma unroll
        for (int k = 0; k < DP; ++k) gz[s][k] = (live[s] && k < d) ? a.g_in[j * d + k] : T(0);
      }
    } else
#pragma unroll
    for (int s = 0; s < S; ++s) {
      T q = 0;
#pragma unroll
      for (int k = 0; k < DP; ++k) {
        T u = 0, is = 1;
        if (k < d) {
          if (a.base) { is = 1 / a.base[d + k]; u = (z[s][k] - a.base[k]) * is; } else u = z[s][k];
        }
        q += u * u;
        gz[s][k] = live[s] ? -u * is : T(0);
      }
      const T term = a.base_c0 - q / 2 + ld[s];
      if (live[s]) {
        obj_local += term;
        if (a.flags & EW_WRITE_TERMS) a.terms_out[gi * group + (int64_t)s * nthr + tid] = term;
      }
    }
    if (!want_grad) continue;
    // ---- backward sweep: layers L-1 .. 0, walking the forward maps from x0 back to y ----
    for (int l = L - 1; l >= 0; --l) {
      const T* e = s_tab + (size_t)l * STR;
      T* acc = s_acc + ((size_t)(tid >> 5) * L + l) * NACC;
      const int kind = s_kind[l];
      if (kind == NF_PLANAR) {
        const T b = e[0], m = e[1];
        T ga[S], tt[S], gyn[S][DP];
        T g_m = 0, g_b = 0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const T w8 = wl[s];
          T dot = b;
#pragma unroll
          for (int k = 0; k < DP; ++k) dot += e[4 + k] * z[s][k];
          const T t = N_::tanh(dot);
          const T psi = 1 - t * t, den = 1 + m * psi;
          // g_hat = gz - grad_z logdet_fwd = gz + (2 t m psi / den) w ;  gy = g_hat - psi w (u_hat . g_hat) / den
          const T cw = w8 * 2 * t * m * psi / den;
          T ug = 0;
#pragma unroll
          for (int k = 0; k < DP; ++k) { gyn[s][k] = gz[s][k] + cw * e[4 + k]; ug += e[4 + DP + k] * gyn[s][k]; }
          const T sc = psi * ug / den;
          T ugy = 0;
#pragma unroll
          for (int k = 0; k < DP; ++k) { gyn[s][k] -= sc * e[4 + k]; ugy += e[4 + DP + k] * gyn[s][k]; }
          // parameter gradients = forward-layer backward at z with upstream (-gy, -1)
          const T gt = -ugy + w8 * 2 * t * m / den;
          ga[s] = psi * gt; tt[s] = t;
          g_m += -w8 * psi / den; g_b += ga[s];
        }
#pragma unroll
        for (int k = 0; k < DP; ++k) {
          T gu = 0, gw = 0;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            gu += -tt[s] * gyn[s][k];
            gw += z[s][k] * ga[s];
          }
          if (k < d) {
            gu = warp_sum(gu); gw = warp_sum(gw);
            if (lane == 0) { acc[k] += gu; acc[DP + k] += gw; }
          }
        }
        g_m = warp_sum(g_m); g_b = warp_sum(g_b);
        if (lane == 0) { acc[2 * DP] += g_m; acc[2 * DP + 1] += g_b; }
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int k = 0; k < DP; ++k) { z[s][k] += e[4 + DP + k] * tt[s]; gz[s][k] = gyn[s][k]; }
      } else if (kind == NF_RADIAL) {
        const T alpha = e[0], bh = e[1];
        T g_al = 0, g_bh = 0, gz0[DP];
#pragma unroll
        for (int k = 0; k < DP; ++k) gz0[k] = 0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const T w8 = wl[s];
          T u[DP], r2 = 0;
#pragma unroll
          for (int k = 0; k < DP; ++k) { u[k] = (k < d) ? z[s][k] - e[4 + k] : T(0); r2 += u[k] * u[k]; }
          const T r = N_::sqrt(r2), ir = r > 0 ? 1 / r : T(0);
          const T h = 1 / (alpha + r), g = bh * h, gp = -bh * h * h;
          const T A = 1 + g, Bq = 1 + alpha * bh * h * h;
          const T dld_dh = T(d - 1) * bh / A + 2 * alpha * bh * h / Bq;
          const T dld_dr = -h * h * dld_dh;
          // g_hat = gz - (dld/dr) u / r ;  J = A I + (gp / r) u u^T  (symmetric) ;  gy = J^{-1} g_hat
          T gh[DP], ugh = 0;
#pragma unroll
          for (int k = 0; k < DP; ++k) { gh[k] = gz[s][k] - w8 * dld_dr * u[k] * ir; ugh += u[k] * gh[k]; }
          const T cj = (gp * ir) / (A + gp * r);
          T gy[DP], ugy = 0;
#pragma unroll
          for (int k = 0; k < DP; ++k) { gy[k] = (gh[k] - cj * u[k] * ugh) / A; ugy += u[k] * gy[k]; }
          // parameter gradients = forward-layer backward at z with upstream (-gy, -1)
          g_al += -ugy * gp - w8 * ((-h * h) * dld_dh + bh * h * h / Bq);
          g_bh += -ugy * h - w8 * (T(d - 1) * h / A + alpha * h * h / Bq);
          // d f / d z0 = -(J - I)  =>  contribution (J - I)^T gy  with upstream -gy, plus the logdet term
          const T coef = -ugy * gp - w8 * dld_dr;       // same r-channel as the forward kernel, upstream (-gy, -1)
#pragma unroll
          for (int k = 0; k < DP; ++k) {
            gz0[k] += g * gy[k] - coef * u[k] * ir;
            z[s][k] = (k < d) ? z[s][k] + g * u[k] : T(0);
            gz[s][k] = gy[k];
          }
        }
#pragma unroll
        for (int k = 0; k < DP; ++k)
          if (k < d) {
            const T v = warp_sum(gz0[k]);
            if (lane == 0) acc[k] += v;
          }
        g_al = warp_sum(g_al); g_bh = warp_sum(g_bh);
        if (lane == 0) { acc[2 * DP] += g_al; acc[2 * DP + 1] += g_bh; }
      } else if (kind == NF_SHIFT) {
#pragma unroll
        for (int k = 0; k < DP; ++k) {
          T v = 0;
#pragma unroll
          for (int s = 0; s < S; ++s) { v -= gz[s][k]; z[s][k] += e[4 + k]; }
          if (k < d) { v = warp_sum(v); if (lane == 0) acc[k] += v; }
        }
      } else if (kind == NF_SCALE) {  // z = y / a
#pragma unroll
        for (int k = 0; k < DP; ++k) {
          T v = 0;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            v -= gz[s][k] * z[s][k] * e[4 + DP + k];
            gz[s][k] *= e[4 + DP + k];
            z[s][k] *= e[4 + k];
          }
          if (k < d) { v = warp_sum(v); if (lane == 0) acc[k] += v; }
        }
      } else if (kind == NF_MOMENTUM_AFFINE) {  // rho_in = (rho_out - b) / a
#pragma unroll
        for (int k = HD; k < DP; ++k) {
          T gb = 0, ga = 0;
          const T ia = 1 / e[4 + DP + k];
#pragma unroll
          for (int s = 0; s < S; ++s) {
            gz[s][k] *= ia;
            gb -= gz[s][k];
            ga -= gz[s][k] * z[s][k];
            z[s][k] = z[s][k] * e[4 + DP + k] + e[4 + k];
          }
          gb = warp_sum(gb); ga = warp_sum(ga);
          if (lane == 0) { acc[k] += gb; acc[DP + k] += ga; }
        }
      } else {  // NF_LEAPFROG (applied with -eps)
        T eps[HD], ge[HD];
#pragma unroll
        for (int k = 0; k < HD; ++k) { eps[k] = e[4 + k]; ge[k] = 0; }
        const int nst = (int)e[0];
#pragma unroll
        for (int s = 0; s < S; ++s)
          leapfrog_backward<T, HD, LR>(a.sp, z[s], z[s] + HD, gz[s], gz[s] + HD, eps, T(-1), nst, ge);
#pragma unroll
        for (int k = 0; k < HD; ++k) {
          const T v = warp_sum(ge[k]);
          if (lane == 0) acc[k] += v;
        }
      }
    }
    if (a.flags & EW_EXT_GRAD) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int64_t j = gi * group + (int64_t)s * nthr + tid;
        if (!live[s]) continue;
#pragma unroll
        for (int k = 0; k < DP; ++k)
          if (k < d) a.g_out[j * d + k] = gz[s][k];
      }
    }
  }
  __syncthreads();
  if (a.gpart)
    for (int i = tid; i < L * NACC; i += nthr) {
      T t = 0;
      for (int w = 0; w < nwarp; ++w) t += s_acc[(size_t)w * L * NACC + i];
      a.gpart[(size_t)blockIdx.x * L * NACC + i] = t;
    }
  if (a.epart) {
    double ev = warp_sum((double)obj_local);
    __shared__ double s_e[32];
    if ((tid & 31) == 0) s_e[tid >> 5] = ev;
    __syncthreads();
    if (tid == 0) {
      double t = 0;
      for (int w = 0; w < (nthr + 31) / 32; ++w) t += s_e[w];
      a.epart[blockIdx.x] = t;
    }
  }
}

// Deterministic cross-CTA reduction + chain rule from the reduced per-layer sums to theta order.
// One thread per layer; gsum[P+1] receives UNSCALED sums (gradient sums then the ELBO sum at [P]).
template <typename T, int DP>
__device__ __forceinline__ void ew_finalize_body(const T* __restrict__ theta, const EwLayerMeta* __restrict__ meta, int L, int d,
                                                 const T* __restrict__ gpart, const double* __restrict__ epart, int nblocks,
                                                 int64_t N, int64_t P, int want_grad, int inverse, double* __restrict__ gsum, int l,
                                                 int accumulate = 0) {
  constexpr int NACC = ew_nacc<DP>();
  if (l == 0 && epart) {
    double e = 0;
    for (int b = 0; b < nblocks; ++b) e += epart[b];
    gsum[P] = e;
  }
  if (l >= L || !want_grad) return;
  double G[NACC];
  for (int i = 0; i < NACC; ++i) G[i] = 0;
  for (int b = 0; b < nblocks; ++b)
    for (int i = 0; i < NACC; ++i) G[i] += (double)gpart[((size_t)b * L + l) * NACC + i];
  const T* p = theta + meta[l].theta_off;
  double* gdst = gsum + meta[l].theta_off;
  double gtmp[2 * DP + 2];
  double* g = accumulate ? gtmp : gdst;        // a segment of a layered flow adds into sums other chunks may already hold
  switch (meta[l].kind) {
    case NF_PLANAR: {  // App. A.1 hand backward through u_hat(u, w)
      const T* w = p; const T* u = p + d;
      double s = 0, n = 0, gk = 0;
      for (int k = 0; k < d; ++k) { s += (double)w[k] * u[k]; n += (double)w[k] * w[k]; gk += (double)w[k] * G[k]; }
      const double m = softplus_stable<double>(s) - 1;
      const double kappa = (m - s) / n;
      const double g_m = G[2 * DP] + gk / n;
      const double g_s = -gk / n + sigmoid_stable<double>(s) * g_m;
      const double g_n = -kappa * gk / n;
      for (int k = 0; k < d; ++k) {
        g[k] = G[DP + k] + kappa * G[k] + g_s * u[k] + 2 * g_n * w[k];   // w
        g[d + k] = G[k] + g_s * w[k];                                     // u
      }
      g[2 * d] = G[2 * DP + 1];                                           // b
      break;
    }
    case NF_RADIAL: {
      const double sa = sigmoid_stable<double>((double)p[0]), sb = sigmoid_stable<double>((double)p[1]);
      g[0] = sa * (G[2 * DP] - G[2 * DP + 1]);
      g[1] = sb * G[2 * DP + 1];
      for (int k = 0; k < d; ++k) g[2 + k] = G[k];
      break;
    }
    case NF_SHIFT:
      for (int k = 0; k < d; ++k) g[k] = G[k];
      break;
    case NF_SCALE:
      for (int k = 0; k < d; ++k) g[k] = G[k] + (inverse ? -1.0 : 1.0) * (double)N / (double)p[k];
      break;
    case NF_MOMENTUM_AFFINE: {
      const int h = d / 2;
      for (int k = 0; k < h; ++k) {
        g[k] = G[h + k];
        g[h + k] = G[DP + h + k] + (inverse ? -1.0 : 1.0) * (double)N / (double)p[h + k];
      }
      break;
    }
    case NF_LEAPFROG: {
      const int h = d / 2;
      for (int k = 0; k < h; ++k) g[k] = G[k] * exp((double)p[k]);
      break;
    }
  }
  if (accumulate) {
    int np = 0;
    switch (meta[l].kind) {
      case NF_PLANAR: np = 2 * d + 1; break;
      case NF_RADIAL: np = d + 2; break;
      case NF_SHIFT: case NF_SCALE: np = d; break;
      case NF_MOMENTUM_AFFINE: np = d; break;
      case NF_LEAPFROG: np = d / 2; break;
    }
    for (int k = 0; k < np; ++k) gdst[k] += gtmp[k];
  }
}

template <typename T, int DP>
__global__ void ew_finalize_kernel(const T* __restrict__ theta, const EwLayerMeta* __restrict__ meta, int L, int d,
                                   const T* __restrict__ gpart, const double* __restrict__ epart, int nblocks,
                                   int64_t N, int64_t P, int want_grad, int inverse, double* __restrict__ gsum, int accumulate = 0) {
  ew_finalize_body<T, DP>(theta, meta, L, d, gpart, epart, nblocks, N, P, want_grad, inverse, gsum, (int)(blockIdx.x * blockDim.x + threadIdx.x),
                          accumulate);
}

// ---------------------------------------------------------------------------------------------
// Persistent training loop for small batches (the reference's own demo regime: planar flow, batch 10-64, 10^4-10^5 Adam
// iterations -- example/demo_planar_flow.jl:25-47).  One CTA runs EVERY iteration inside a single launch: layer-table prep,
// fused forward + target + backward over the batch (device Philox draws, seed + iteration as in the multi-launch loop),
// chain rule to theta, Optimisers.Adam step, and the (loss, |g|^2) record of reference src/optimize.jl:89 -- phases separated by
// __syncthreads instead of kernel boundaries, so an iteration costs a few microseconds instead of four launches.
// ---------------------------------------------------------------------------------------------
template <typename T> struct EwTrain {
  T* theta; T* m; T* v;            // [P] device
  const EwLayerMeta* meta;
  T* table;                        // [L, stride]
  double* gsum;                    // [P+1]
  double* stats;                   // [n_iters][2]: loss, |g|^2
  int64_t P;
  int n_iters, t0;
  double eta, b1, b2, eps;
  uint64_t seed0;
};

template <typename T, int DP, int S, bool LR>
__global__ void __launch_bounds__(128) ew_train_kernel(EwArgs<T> a, EwTrain<T> tr) {
  __shared__ double s_g2[4];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const double inv_n = 1.0 / (double)a.N;
  double b1t = pow(tr.b1, (double)tr.t0), b2t = pow(tr.b2, (double)tr.t0);     // beta^t, advanced by one product per iteration
  for (int it = 0; it < tr.n_iters; ++it) {
    for (int l = tid; l < a.L; l += nthr) ew_prep_body<T, DP>(tr.theta, tr.meta, l, a.d, tr.table);
    __syncthreads();
    EwArgs<T> ai = a;
    ai.seed = tr.seed0 + (uint64_t)it;
    ew_flow_body<T, DP, S, LR>(ai, 0, 1);
    __syncthreads();
    for (int l = tid; l < a.L; l += nthr)
      ew_finalize_body<T, DP>(tr.theta, tr.meta, a.L, a.d, a.gpart, a.epart, 1, a.N, tr.P, 1, 0, tr.gsum, l);
    __syncthreads();
    const T b1 = (T)tr.b1, b2 = (T)tr.b2, eta = (T)tr.eta, eps = (T)tr.eps;
    b1t *= tr.b1; b2t *= tr.b2;
    const T omb1t = (T)(1.0 - b1t), omb2t = (T)(1.0 - b2t);
    double g2 = 0;
    for (int64_t i = tid; i < tr.P; i += nthr) {
      const T g = (T)(-tr.gsum[i] * inv_n);
      const T mi = b1 * tr.m[i] + (T(1) - b1) * g;
      const T vi = b2 * tr.v[i] + (T(1) - b2) * g * g;
      tr.m[i] = mi; tr.v[i] = vi;
      tr.theta[i] -= mi / omb1t / (Num<T>::sqrt(vi / omb2t) + eps) * eta;
      g2 += (double)g * (double)g;
    }
    g2 = warp_sum(g2);
    if ((tid & 31) == 0) s_g2[tid >> 5] = g2;
    __syncthreads();
    if (tid == 0) {
      double t = 0;
      for (int w = 0; w < (nthr + 31) / 32; ++w) t += s_g2[w];
      tr.stats[2 * it] = -tr.gsum[tr.P] * inv_n;
      tr.stats[2 * it + 1] = t;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
template <typename T, int DP, int S, bool LR, bool INV>
static int ew_launch(Flow& f, const Target* tgt, const T* theta_dev, int64_t N, const T* z0_dev, uint64_t seed,
                     int flags, T* y_out, T* ld_out, T* terms_out, double* gsum_dev, bool inverse) {
  const int L = (int)f.layers.size(), d = f.dim;
  constexpr int STR = ew_stride<DP>(), NACC = ew_nacc<DP>();
  const int threads = 128;
  const size_t smem = ((size_t)L * STR + (size_t)(threads / 32) * L * NACC + (inverse ? 0 : (size_t)L * S * threads)) * sizeof(T) + (size_t)L * sizeof(int) + 16;
  if (smem > 200 * 1024) {
    set_error("elementwise flow with %d layers needs %zu B of shared memory (limit 200 KiB)", L, smem);
    return NF_ERR_UNSUPPORTED;
  }
  (void)inverse;
  void (*kern)(EwArgs<T>);
  if constexpr (INV) kern = ew_inv_flow_kernel<T, DP, S, LR>; else kern = ew_flow_kernel<T, DP, S, LR>;
  NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t group = (int64_t)threads * S;
  int max_blocks = 0;
  NF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, kern, threads, smem));
  if (max_blocks < 1) max_blocks = 1;
  const int grid = (int)std::min<int64_t>(ceil_div(N, group), (int64_t)kNumSMs * max_blocks);

  T* table = (T*)f.ws_alloc((size_t)L * STR * sizeof(T));
  T* gpart = (T*)f.ws_alloc((size_t)grid * L * NACC * sizeof(T));
  double* epart = (double*)f.ws_alloc((size_t)grid * sizeof(double));
  if (!table || !gpart || !epart) return NF_ERR_OOM;

  ew_prep_kernel<T, DP><<<(L + 63) / 64, 64, 0, f.stream>>>(theta_dev, f.d_ew_meta, L, d, table);
  NF_LAUNCH_CHECK();
  EwArgs<T> a{};
  a.z0 = z0_dev; a.table = table; a.kinds = f.d_ew_kinds;
  a.base = f.base_is_standard ? nullptr : (const T*)f.d_base;
  a.base_c0 = (T)f.base_c0;
  if (f.base_dense) {
    // q0 = MvNormal(mu, L L^T): the draws (or the caller's x0) and their log-density come from base_dense_kernel; the fused
    // kernel then sees ready-made x0 and a per-sample log q0
    if (inverse) {
      set_error("full-covariance base: the inverse direction of elementwise flows runs on the layered path (internal routing error)");
      return NF_ERR_UNSUPPORTED;
    }
    T* lq0 = (T*)f.ws_alloc((size_t)N * sizeof(T));
    if (!lq0) return NF_ERR_OOM;
    if (!z0_dev) {
      T* x0 = (T*)f.ws_alloc((size_t)N * d * sizeof(T));
      if (!x0) return NF_ERR_OOM;
      base_sample_kernel<T><<<base_sample_grid(N, d, f.draw_row_offset), 256, 0, f.stream>>>(x0, nullptr, d, N, seed, f.draw_row_offset);
      NF_LAUNCH_CHECK();
      NF_TRY(base_dense_launch<T>(f, x0, N, 0, lq0, nullptr));
      z0_dev = x0;
    } else {
      NF_TRY(base_dense_launch<T>(f, const_cast<T*>(z0_dev), N, 1, lq0, nullptr));
    }
    a.z0 = z0_dev; a.base = nullptr; a.lq0 = lq0;
  }
  if (tgt) a.tp = tgt->params<T>();
  if (f.score_target) a.sp = f.score_target->params<T>();
  a.y_out = y_out; a.ld_out = ld_out; a.terms_out = terms_out;
  a.gpart = gpart; a.epart = epart; a.N = N; a.L = L; a.d = d;
  a.flags = flags | (z0_dev ? 0 : EW_GEN_Z0);
  a.seed = seed;
  a.row0 = f.draw_row_offset;
  f.prof.begin("ew_flow", f.stream);
  kern<<<grid, threads, smem, f.stream>>>(a);
  f.prof.end(f.stream);
  NF_LAUNCH_CHECK();
  if (gsum_dev) {
    ew_finalize_kernel<T, DP><<<(L + 63) / 64, 64, 0, f.stream>>>(theta_dev, f.d_ew_meta, L, d, gpart, epart, grid, N,
                                                                 f.P, (flags & EW_GRAD) ? 1 : 0, inverse ? 1 : 0, gsum_dev);
    NF_LAUNCH_CHECK();
  }
  return NF_OK;
}

// A run of elementwise layers [l0, l0 + Lseg) INSIDE a layered (coupling) flow, forward direction or (INV) the inverse direction (reference src/flows/utils.jl:23-26: any
// composition of bijectors is a flow).  forward: Xout = T_seg(Xin), ld += logdet.  backward: recomputes the run from its input
// state Xin, pulls G (d/dy -> d/dx, in place) and the per-sample logdet weights gld (nullptr = 1) back through it and ADDS the
// parameter-gradient sums into gsum.
template <typename T, int DP, int S, bool INV>
static int ew_segment_launch(Flow& f, int l0, int Lseg, const T* theta_dev, int64_t N, const T* Xin, T* Xout, T* ld, bool backward,
                             T* G, const T* gld, double* gsum) {
  const int d = f.dim;
  constexpr int STR = ew_stride<DP>(), NACC = ew_nacc<DP>();
  const int threads = 128;
  const size_t smem = ((size_t)Lseg * STR + (size_t)(threads / 32) * Lseg * NACC + (INV ? 0 : (size_t)Lseg * S * threads)) * sizeof(T) + (size_t)Lseg * sizeof(int) + 16;
  if (smem > 200 * 1024) {
    set_error("a run of %d elementwise layers needs %zu B of shared memory (limit 200 KiB)", Lseg, smem);
    return NF_ERR_UNSUPPORTED;
  }
  void (*kern)(EwArgs<T>);
  if constexpr (INV) kern = ew_inv_flow_kernel<T, DP, S, false>; else kern = ew_flow_kernel<T, DP, S, false>;
  NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t group = (int64_t)threads * S;
  int max_blocks = 0;
  NF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, kern, threads, smem));
  if (max_blocks < 1) max_blocks = 1;
  const int grid = (int)std::min<int64_t>(ceil_div(N, group), (int64_t)kNumSMs * max_blocks);
  T* table = (T*)f.ws_alloc((size_t)Lseg * STR * sizeof(T));
  T* gpart = backward ? (T*)f.ws_alloc((size_t)grid * Lseg * NACC * sizeof(T)) : nullptr;
  if (!table || (backward && !gpart)) return NF_ERR_OOM;
  ew_prep_kernel<T, DP><<<(Lseg + 63) / 64, 64, 0, f.stream>>>(theta_dev, f.d_ew_meta + l0, Lseg, d, table);
  NF_LAUNCH_CHECK();
  EwArgs<T> a{};
  a.z0 = Xin; a.table = table; a.kinds = f.d_ew_kinds + l0;
  a.N = N; a.L = Lseg; a.d = d;
  if (!backward) { a.flags = EW_WRITE_Y | EW_ADD_LD; a.y_out = Xout; a.ld_out = ld; }
  else { a.flags = EW_GRAD | EW_EXT_GRAD; a.g_in = G; a.g_out = G; a.gld = gld; a.gpart = gpart; }
  f.prof.begin("ew_segment", f.stream);
  kern<<<grid, threads, smem, f.stream>>>(a);
  f.prof.end(f.stream);
  NF_LAUNCH_CHECK();
  if (backward) {
    ew_finalize_kernel<T, DP><<<(Lseg + 63) / 64, 64, 0, f.stream>>>(theta_dev, f.d_ew_meta + l0, Lseg, d, gpart, nullptr, grid, N, f.P, 1, INV ? 1 : 0, gsum, 1);
    NF_LAUNCH_CHECK();
  }
  return NF_OK;
}

template <typename T, bool INV>
int ew_segment_dir(Flow& f, int l0, int Lseg, const void* theta_dev, int64_t N, const void* Xin, void* Xout, void* ld, bool backward,
                   void* G, const void* gld, double* gsum) {
  const int d = f.dim;
#define NF_SEG(DPV, SV) return ew_segment_launch<T, DPV, SV, INV>(f, l0, Lseg, (const T*)theta_dev, N, (const T*)Xin, (T*)Xout, (T*)ld, backward, (T*)G, (const T*)gld, gsum)
  if (d <= 2) NF_SEG(2, 4);
  if (d <= 4) NF_SEG(4, 2);
  if (d <= 8) NF_SEG(8, 1);
  if (d <= 16) NF_SEG(16, 1);
  if (d <= 32) NF_SEG(32, 1);
  if (d <= 64) NF_SEG(64, 1);
#undef NF_SEG
  set_error("planar / radial layers inside a coupling flow support dim <= 64, got %d", d);
  return NF_ERR_UNSUPPORTED;
}

// One direction (forward sweep / inverse sweep) per translation unit: the kernels are many large instantiations, and
// splitting them (elementwise_{fwd,inv,train}_{f32,f64}.cu) lets the build compile them in parallel.
template <typename T, bool INV>
int ew_run_dir(Flow& f, const Target* tgt, const void* theta_dev, int64_t N, const void* z0_dev, uint64_t seed,
               bool want_grad, void* y_out, void* ld_out, void* terms_out, double* gsum_dev, bool head) {
  constexpr bool inverse = INV;
  int flags = 0;
  if (tgt || (inverse && head)) flags |= EW_TARGET;
  if (want_grad) flags |= EW_GRAD;
  if (y_out) flags |= EW_WRITE_Y;
  if (ld_out) flags |= EW_WRITE_LD;
  if (terms_out) flags |= EW_WRITE_TERMS;
  const int d = f.dim;
  // the logistic-regression target (as the objective's target or as the LeapFrog score) gets its own instantiations
  const bool lr = (tgt && tgt->kind == NF_TARGET_LOGREG) || (f.score_target && f.score_target->kind == NF_TARGET_LOGREG);
  // Hamiltonian flows whose state does not fit one thread's registers (BASELINE config 5: 100-D posterior): a warp per sample
  if (hmc_warp_qualifies(f, INV ? nullptr : tgt) && (g_opt_hmc_warp || d > 64 || (lr && d > 16) || (d & (d - 1)) != 0)) {
    if constexpr (INV) return hmc_warp_inverse<T>(f, theta_dev, N, z0_dev, head, want_grad, y_out, ld_out, terms_out, gsum_dev);
    else return hmc_warp_run<T>(f, tgt, theta_dev, N, z0_dev, seed, want_grad, y_out, ld_out, terms_out, gsum_dev);
  }
  if ((f.hamiltonian || (tgt && tgt->joint)) && (d & (d - 1)) != 0) {
    set_error("Hamiltonian flows / joint targets need dim a power of two on this path (inverse direction, or layers / targets the "
              "warp-per-sample kernel does not cover), got %d", d);
    return NF_ERR_UNSUPPORTED;
  }
  if (lr && d > 16) {
    set_error("elementwise flows on the logistic-regression target support dim <= 16, got %d", d);
    return NF_ERR_UNSUPPORTED;
  }
#define NF_EW_ARGS f, tgt, (const T*)theta_dev, N, (const T*)z0_dev, seed, flags, (T*)y_out, (T*)ld_out, (T*)terms_out, gsum_dev, inverse
#define NF_EW_CASE(DPV, SV)                                                      \
  do {                                                                           \
    if (lr) { if constexpr (DPV <= 16) return ew_launch<T, DPV, SV, true, INV>(NF_EW_ARGS); } \
    return ew_launch<T, DPV, SV, false, INV>(NF_EW_ARGS);                        \
  } while (0)
  if (d <= 2) NF_EW_CASE(2, 4);
  if (d <= 4) NF_EW_CASE(4, 2);
  if (d <= 8) NF_EW_CASE(8, 1);
  if (d <= 16) NF_EW_CASE(16, 1);
  if (d <= 32) NF_EW_CASE(32, 1);
  if (d <= 64) NF_EW_CASE(64, 1);
#undef NF_EW_CASE
#undef NF_EW_ARGS
  set_error("elementwise (planar/radial) flows support dim <= 64 in this build, got %d", d);
  return NF_ERR_UNSUPPORTED;
}

template <typename T, int DP, int S, bool LR>
static int ew_train_launch(Flow& f, const Target* tgt, int64_t N, uint64_t seed, int n_iters, int t0, double eta, double b1,
                           double b2, double eps, void* m_dev, void* v_dev) {
  const int L = (int)f.layers.size(), d = f.dim;
  constexpr int STR = ew_stride<DP>(), NACC = ew_nacc<DP>();
  const int threads = 128;
  const size_t smem = ((size_t)L * STR + (size_t)(threads / 32) * L * NACC + (size_t)L * S * threads) * sizeof(T) + (size_t)L * sizeof(int) + 16;
  if (smem > 200 * 1024) {
    set_error("elementwise flow with %d layers needs %zu B of shared memory (limit 200 KiB)", L, smem);
    return NF_ERR_UNSUPPORTED;
  }
  auto kern = ew_train_kernel<T, DP, S, LR>;
  NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  f.ws_reset();
  NF_TRY(f.ws_reserve((size_t)4 << 20));
  T* table = (T*)f.ws_alloc((size_t)L * STR * sizeof(T));
  T* gpart = (T*)f.ws_alloc((size_t)L * NACC * sizeof(T));
  double* epart = (double*)f.ws_alloc(sizeof(double));
  if (!table || !gpart || !epart) return NF_ERR_OOM;
  EwArgs<T> a{};
  a.z0 = nullptr; a.table = table; a.kinds = f.d_ew_kinds;
  a.base = f.base_is_standard ? nullptr : (const T*)f.d_base;
  a.base_c0 = (T)f.base_c0;
  a.tp = tgt->params<T>();
  if (f.score_target) a.sp = f.score_target->params<T>();
  a.gpart = gpart; a.epart = epart; a.N = N; a.L = L; a.d = d;
  a.flags = EW_GRAD | EW_TARGET | EW_GEN_Z0;
  EwTrain<T> tr{};
  tr.theta = (T*)f.d_theta; tr.m = (T*)m_dev; tr.v = (T*)v_dev; tr.meta = f.d_ew_meta; tr.table = table; tr.gsum = f.d_gsum;
  tr.stats = f.d_stats; tr.P = f.P; tr.n_iters = n_iters; tr.t0 = t0; tr.eta = eta; tr.b1 = b1; tr.b2 = b2; tr.eps = eps; tr.seed0 = seed;
  f.prof.begin("ew_train", f.stream);
  kern<<<1, threads, smem, f.stream>>>(a, tr);
  f.prof.end(f.stream);
  NF_LAUNCH_CHECK();
  return NF_OK;
}

// n_iters Adam iterations of the ELBO objective in ONE launch (single CTA); for batches small enough that one CTA is the
// right amount of hardware (see ew_train_kernel).  Returns NF_ERR_UNSUPPORTED when the flow / batch does not qualify.
template <typename T>
int ew_train(Flow& f, const Target* tgt, int64_t N, uint64_t seed, int n_iters, int t0, double eta, double b1, double b2,
             double eps, void* m_dev, void* v_dev) {
  const int d = f.dim;
  if ((f.hamiltonian || tgt->joint) && (d & (d - 1)) != 0) {
    set_error("Hamiltonian flows / joint targets need dim a power of two, got %d", d);
    return NF_ERR_UNSUPPORTED;
  }
  const bool lr = tgt->kind == NF_TARGET_LOGREG || (f.score_target && f.score_target->kind == NF_TARGET_LOGREG);
#define NF_EWT_CASE(DPV, SV)                                                                                              \
  do {                                                                                                                    \
    if (lr) return ew_train_launch<T, DPV, SV, true>(f, tgt, N, seed, n_iters, t0, eta, b1, b2, eps, m_dev, v_dev);        \
    return ew_train_launch<T, DPV, SV, false>(f, tgt, N, seed, n_iters, t0, eta, b1, b2, eps, m_dev, v_dev);             \
  } while (0)
  // one sample per thread while the batch fits the CTA (a thread's S samples are a serial dependency chain)
  if (d <= 2) { if (N <= 128) NF_EWT_CASE(2, 1); NF_EWT_CASE(2, 4); }
  if (d <= 4) { if (N <= 128) NF_EWT_CASE(4, 1); NF_EWT_CASE(4, 2); }
  if (d <= 8) NF_EWT_CASE(8, 1);
  if (d <= 16) NF_EWT_CASE(16, 1);
#undef NF_EWT_CASE
  set_error("persistent training kernel supports dim <= 16, got %d", d);
  return NF_ERR_UNSUPPORTED;
}

}  // namespace nf
