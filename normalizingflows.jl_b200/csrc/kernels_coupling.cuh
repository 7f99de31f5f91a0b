// Device kernels of the layered coupling path: CUDA-core GEMMs with fused epilogues (the Float64 /
// validation mainloop; Float32 flows use the tcgen05 mainloop of tc_gemm.cu), the affine-coupling
// and rational-quadratic-spline elementwise kernels, target / base kernels.
//
// Reference rows: a8 fnn/Dense (src/flows/utils.jl:71-100), a11 AffineCoupling
// (src/flows/realnvp.jl:57-110), a12/a13 NeuralSplineCoupling + MonotonicSplines RQS
// (src/flows/neuralspline.jl:65-140; App. A.4), a14 PartitionMask, a15 base logpdf, a16 targets.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "targets.cuh"

namespace nf {

// ---------------------------------------------------------------------------------------------
// epilogues of the CUDA-core GEMM:  C(m, n) = sum_k A(m, k) B(k, n)
// ---------------------------------------------------------------------------------------------
enum : int { ACT_NONE = 0, ACT_LRELU = 1, ACT_TANH = 2 };

template <typename T> struct EpiBiasAct {       // Dense forward: act(W x + b)
  T* C; int ldc; const T* bias; int act;
  __device__ __forceinline__ void operator()(int64_t m, int n, T v) const {
    v += bias[n];
    if (act == ACT_LRELU) v = v > 0 ? v : T(0.01) * v;
    else if (act == ACT_TANH) v = Num<T>::tanh(v);
    C[m * ldc + n] = v;
  }
};
template <typename T> struct EpiMaskLrelu {     // dgrad into a hidden layer: multiply by leakyrelu'(h)
  T* C; int ldc; const T* ref; int ldr;
  __device__ __forceinline__ void operator()(int64_t m, int n, T v) const {
    C[m * ldc + n] = ref[m * ldr + n] > 0 ? v : T(0.01) * v;
  }
};
template <typename T> struct EpiScatterAdd {    // dgrad into the conditioner input: G[:, idx2] += .
  T* G; int ldg; const int* idx;
  __device__ __forceinline__ void operator()(int64_t m, int n, T v) const { G[m * ldg + idx[n]] += v; }
};
template <typename T> struct EpiAtomicDouble {  // wgrad partial -> double accumulators in theta order
  double* g; int ldg;
  __device__ __forceinline__ void operator()(int64_t m, int n, T v) const { atomicAdd(&g[m * ldg + n], (double)v); }
};

template <typename T, bool TA, bool TB, typename Epi>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ B,
                                                        int64_t ldb, int64_t M, int N, int64_t K, int64_t ksplit, Epi epi) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ T As[BK][BM + 4];
  __shared__ T Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * ksplit;
  const int64_t kend = kbeg + ksplit < K ? kbeg + ksplit : K;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int mm, kk;
      if (TA) { mm = e & 63; kk = e >> 6; } else { kk = e & 15; mm = e >> 4; }
      const int64_t gm = m0 + mm, gk = k0 + kk;
      T v = 0;
      if (gm < M && gk < kend) v = TA ? A[gk * lda + gm] : A[gm * lda + gk];
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int nn, kk;
      if (TB) { kk = e & 15; nn = e >> 4; } else { nn = e & 63; kk = e >> 6; }
      const int gn = n0 + nn;
      const int64_t gk = k0 + kk;
      T v = 0;
      if (gn < N && gk < kend) v = TB ? B[(int64_t)gn * ldb + gk] : B[gk * ldb + gn];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t m = m0 + ty * 4 + i;
      const int n = n0 + tx * 4 + j;
      if (m < M && n < N) epi(m, n, acc[i][j]);
    }
}

// column sums of g [N x n] -> double atomics (bias gradients).  blockDim = 256 = 32 column lanes x 8 row lanes.
template <typename T>
__global__ void colsum_atomic_kernel(const T* __restrict__ g, int64_t N, int n, int64_t rows_per_block, double* __restrict__ out) {
  __shared__ double red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < N ? r0 + rows_per_block : N;
  for (int c0 = 0; c0 < n; c0 += 32) {
    const int c = c0 + cx;
    double s = 0;
    if (c < n)
      for (int64_t r = r0 + ry; r < r1; r += 8) s += (double)g[r * n + c];
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && c < n) {
      double t = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += red[q][cx];
      atomicAdd(&out[c], t);
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void gather_cols_kernel(const T* __restrict__ X, int d, const int* __restrict__ idx, int n, int64_t N, T* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * n) return;
  const int64_t r = e / n;
  const int k = (int)(e - r * n);
  out[e] = X[r * d + idx[k]];
}

// ---------------------------------------------------------------------------------------------
// AffineCoupling (reference src/flows/realnvp.jl:57-110)
//   forward : y1 = exp(s) .* x1 .+ t,  logjac = sum(s)            (:77-83)
//   inverse : x1 = (y1 .- t) .* exp.(-s), logjac = -sum(s)        (:99-110)
// S holds s AFTER the tanh output activation (|s| < 1), Tt holds t.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// amax_meta[1] holds the bits of a running max |x| (non-negative floats order like unsigned ints).
// Block-level reduction first: ONE atomic per CTA (same-address atomics serialise in L2).
__device__ __forceinline__ void amax_update(float* amax_meta, float v) {
  __shared__ float s_amax[32];
  v = warp_max_f(v);
  if ((threadIdx.x & 31) == 0) s_amax[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float m = threadIdx.x < (blockDim.x >> 5) ? s_amax[threadIdx.x] : 0.f;
    m = warp_max_f(m);
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned int*>(amax_meta + 1), __float_as_uint(m));
  }
  __syncthreads();
}

// One thread per element of the state (coalesced); pos[j] = position of column j in the transformed block or -1.
// Optionally records max |Xout| (exact) for the fp16 operand scaling of the next coupling's conditioner input.
template <typename T, bool INV>
__global__ void affine_apply_kernel(const T* __restrict__ Xin, const T* __restrict__ S, const T* __restrict__ Tt,
                                    const int* __restrict__ pos, int c, int d, int64_t N, T* __restrict__ Xout,
                                    T* __restrict__ ld, float* __restrict__ amax_meta) {
  constexpr int U = 4;                                       // independent elements in flight per thread
  const int64_t total = N * d;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t iters = (total + stride * U - 1) / (stride * U);   // same trip count for every thread (warp collectives inside)
  const bool small = total < ((int64_t)1 << 31);            // 32-bit index math (64-bit division is ~10x the cost)
  float run_max = 0.f;
  for (int64_t it = 0; it < iters; ++it) {
    int64_t e[U], r[U];
    int k[U];
    bool valid[U];
    T x[U], sv[U], tv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      e[u] = (it * U + u) * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      valid[u] = e[u] < total;
      r[u] = 0; k[u] = -1; x[u] = 0;
      if (valid[u]) {
        int j;
        if (small) { const unsigned int r32 = (unsigned int)e[u] / (unsigned int)d; r[u] = r32; j = (int)((unsigned int)e[u] - r32 * (unsigned int)d); }
        else { r[u] = e[u] / d; j = (int)(e[u] - r[u] * d); }
        x[u] = Xin[e[u]];
        k[u] = pos[j];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      sv[u] = 0; tv[u] = 0;
      if (k[u] >= 0) { sv[u] = S[r[u] * c + k[u]]; tv[u] = Tt[r[u] * c + k[u]]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      T y = x[u], contrib = 0;
      if (k[u] >= 0) {
        y = INV ? (x[u] - tv[u]) * Num<T>::exp(-sv[u]) : Num<T>::exp(sv[u]) * x[u] + tv[u];
        contrib = INV ? -sv[u] : sv[u];
      }
      if (valid[u]) {
        Xout[e[u]] = y;
        run_max = fmaxf(run_max, fabsf((float)y));
      }
      if (ld) {
        if ((d & 31) == 0) {            // a warp covers 32 consecutive columns of ONE row
          const T sum = warp_sum(contrib);
          if ((threadIdx.x & 31) == 0 && valid[u]) atomicAdd(&ld[r[u]], sum);
        } else if (valid[u] && k[u] >= 0) {
          atomicAdd(&ld[r[u]], contrib);
        }
      }
    }
  }
  if (amax_meta) amax_update(amax_meta, run_max);
}

// Row-oriented version for d <= 256 (every practical coupling flow): a warp owns R rows at a time, lane l handles columns
// l, l+32, ... so the mask lookup is loop-invariant, every access is a contiguous 128-byte segment, the logdet of a row is
// ONE warp reduction and a plain (non-atomic) update, and there is no per-element index arithmetic.
template <typename T, bool INV, int JMAX>
__global__ void __launch_bounds__(256, 3) affine_apply_rows_kernel(const T* __restrict__ Xin, const T* __restrict__ S,
                                                                const T* __restrict__ Tt, const int* __restrict__ pos, int c, int d,
                                                                int64_t N, T* __restrict__ Xout, T* __restrict__ ld,
                                                                float* __restrict__ amax_meta) {
  constexpr int R0 = JMAX <= 2 ? 4 : (JMAX <= 4 ? 2 : 1);
  constexpr int R = (sizeof(T) == 8 && R0 > 1) ? R0 / 2 : R0;   // rows in flight per warp
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int kk[JMAX];
#pragma unroll
  for (int jj = 0; jj < JMAX; ++jj) { const int j = lane + 32 * jj; kk[jj] = j < d ? pos[j] : -2; }
  float run_max = 0.f;
  for (int64_t r0 = warp * R; r0 < N; r0 += nwarps * R) {
    T x[R][JMAX], sv[R][JMAX], tv[R][JMAX];
    // lane q < R owns row q's logdet slot: fetched with the other loads so the update is not a second HBM round trip
    const bool ld_owner = ld && lane < R && r0 + lane < N;
    const T ldv = ld_owner ? ld[r0 + lane] : T(0);
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int64_t r = r0 + q;
#pragma unroll
      for (int jj = 0; jj < JMAX; ++jj) {
        x[q][jj] = 0; sv[q][jj] = 0; tv[q][jj] = 0;
        if (r < N && kk[jj] >= -1) x[q][jj] = Xin[r * d + lane + 32 * jj];
        if (r < N && kk[jj] >= 0) { sv[q][jj] = S[r * c + kk[jj]]; tv[q][jj] = Tt[r * c + kk[jj]]; }
      }
    }
    T sum[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int64_t r = r0 + q;
      sum[q] = 0;
#pragma unroll
      for (int jj = 0; jj < JMAX; ++jj) {
        T y = x[q][jj];
        if (kk[jj] >= 0) {
          y = INV ? (x[q][jj] - tv[q][jj]) * Num<T>::exp(-sv[q][jj]) : Num<T>::exp(sv[q][jj]) * x[q][jj] + tv[q][jj];
          sum[q] += INV ? -sv[q][jj] : sv[q][jj];
        }
        if (r < N && kk[jj] >= -1) {
          Xout[r * d + lane + 32 * jj] = y;
          run_max = fmaxf(run_max, fabsf((float)y));
        }
      }
    }
    if (ld) {
#pragma unroll
      for (int q = 0; q < R; ++q) sum[q] = warp_sum(sum[q]);
      T mine = sum[0];
#pragma unroll
      for (int q = 1; q < R; ++q) if (lane == q) mine = sum[q];
      if (ld_owner) ld[r0 + lane] = ldv + mine;
    }
  }
  if (amax_meta) amax_update(amax_meta, run_max);
}

// Backward of the coupling arithmetic.  In: G = d/dXout (in place -> d/dXin on idx1 columns; the idx2
// columns receive the conditioner contribution later), gld = d/dlogdet per sample (nullptr -> 1).
// Out: gS = gradient w.r.t. the PRE-tanh output of the s network, gT = gradient w.r.t. t.
// X1src: Xin for the forward direction (x1), Xout for the inverse direction (x1 = output).
template <typename T, bool INV>
__global__ void affine_bwd_kernel(T* __restrict__ G, const T* __restrict__ X1src, const T* __restrict__ S,
                                  const T* __restrict__ gld, const int* __restrict__ idx1, int c, int d, int64_t N,
                                  T* __restrict__ gS, T* __restrict__ gT, float* __restrict__ amaxS, float* __restrict__ amaxT) {
  const int64_t total = N * c;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float maxS = 0.f, maxT = 0.f;
  const bool small = total < ((int64_t)1 << 31);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    T gs_out = 0, gt = 0;
    int64_t r; int k;
    if (small) { const unsigned int r32 = (unsigned int)e / (unsigned int)c; r = r32; k = (int)((unsigned int)e - r32 * (unsigned int)c); }
    else { r = e / c; k = (int)(e - r * c); }
    const int j = idx1[k];
    const T s = S[e];
    const T go = G[r * d + j];
    const T x1 = X1src[r * d + j];
    const T gl = gld ? gld[r] : T(1);
    T gs, gi;
    if (!INV) {
      const T es = Num<T>::exp(s);
      gi = go * es; gs = go * x1 * es + gl; gt = go;
    } else {
      const T ems = Num<T>::exp(-s);
      gi = go * ems; gt = -gi; gs = -go * x1 - gl;
    }
    G[r * d + j] = gi;
    gs_out = gs * (1 - s * s);   // through tanh
    gS[e] = gs_out;
    gT[e] = gt;
    maxS = fmaxf(maxS, fabsf((float)gs_out));
    maxT = fmaxf(maxT, fabsf((float)gt));
  }
  if (amaxS) amax_update(amaxS, maxS);
  if (amaxT) amax_update(amaxT, maxT);
}

// ---------------------------------------------------------------------------------------------
// Bijectors.Shift / Bijectors.Scale inside a layered (coupling) flow: y = x + a / y = a .* x, logdet 0 / sum log|a|
// (e.g. a trainable `Shift ∘ Scale` pre-conditioner in front of RealNVP layers, the pattern of
// reference example/demo_hamiltonian_flow.jl:139-142).  One thread per element; exact max |y| for the next plane scale.
// ---------------------------------------------------------------------------------------------
template <typename T, bool INV>
__global__ void diag_apply_kernel(const T* __restrict__ Xin, const T* __restrict__ a, int is_scale, int d, int64_t N,
                                  T* __restrict__ Xout, T* __restrict__ ld, float* __restrict__ amax_meta) {
  const int64_t total = N * d;
  float run_max = 0.f;
  T slog = 0;
  if (is_scale && ld)
    for (int k = 0; k < d; ++k) slog += Num<T>::log(Num<T>::abs(a[k]));
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / d;
    const int k = (int)(e - r * d);
    const T x = Xin[e];
    const T y = is_scale ? (INV ? x / a[k] : x * a[k]) : (INV ? x - a[k] : x + a[k]);
    Xout[e] = y;
    run_max = fmaxf(run_max, fabsf((float)y));
    if (is_scale && ld && k == 0) ld[r] += INV ? -slog : slog;
  }
  if (amax_meta) amax_update(amax_meta, run_max);
}

// Backward: G (d/dXout) -> d/dXin in place; parameter gradient sums into ga (double atomics).  blockDim = 32 columns x 8 rows.
// V = the values the scale gradient multiplies: the layer input x (forward direction) or the inverse's output x = y / a.
template <typename T, bool INV>
__global__ void diag_bwd_kernel(T* __restrict__ G, const T* __restrict__ V, const T* __restrict__ a, const T* __restrict__ gld,
                                int is_scale, int d, int64_t N, int64_t rows_per_block, double* __restrict__ ga) {
  __shared__ double red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < N ? r0 + rows_per_block : N;
  for (int c0 = 0; c0 < d; c0 += 32) {
    const int c = c0 + cx;
    double s = 0;
    if (c < d) {
      const T ak = a[c];
      for (int64_t r = r0 + ry; r < r1; r += 8) {
        const T g = G[r * d + c];
        const T gl = gld ? gld[r] : T(1);
        if (!is_scale) s += INV ? -(double)g : (double)g;
        else if (!INV) { s += (double)(g * V[r * d + c] + gl / ak); G[r * d + c] = g * ak; }
        else { s += (double)(-g * V[r * d + c] / ak - gl / ak); G[r * d + c] = g / ak; }
      }
    }
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && c < d) {
      double t = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += red[q][cx];
      atomicAdd(&ga[c], t);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Rational-quadratic spline coupling (MonotonicSplines v0.3.3 semantics)
// theta_raw row layout per sample: coordinate i owns rows [i*(3K-1), (i+1)*(3K-1)) = K width logits,
// K height logits, K-1 derivative logits ("block" layout; see oracle RQS_PARAM_LAYOUT).
// Knots: sequential left-to-right softmax sum and cumsum, knot = (2B)*cumsum - B with no FMA contraction.
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
template <typename T> __device__ __forceinline__ T div_rn(T a, T b);
template <> __device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
template <> __device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

template <typename T> struct RqsBin {
  int k;            // searchsortedfirst - 1 : number of knots strictly below v; inside iff 1 <= k <= K
  T x0, x1, y0, y1, d0, d1;
  T cx0, cx1, cy0, cy1;   // cumulative softmax mass at the two knots (for the softmax backward)
  T Sw, Sh;               // softmax denominators
};

template <typename T>
__device__ __forceinline__ void rqs_eval_fwd(const RqsBin<T>& b, T x, T& y, T& lj) {
  using N = Num<T>;
  const T dx = b.x1 - b.x0, dy = b.y1 - b.y0;
  const T s = dy / dx;
  const T xi = (x - b.x0) / dx, om = 1 - xi;
  const T den = s + (b.d1 + b.d0 - 2 * s) * xi * om;
  y = b.y0 + dy * (s * xi * xi + b.d0 * xi * om) / den;
  lj = N::log(N::abs(s * s * (b.d1 * xi * xi + 2 * s * xi * om + b.d0 * om * om))) - 2 * N::log(N::abs(den));
}

template <typename T>
__device__ __forceinline__ void rqs_eval_inv(const RqsBin<T>& b, T y, T& x, T& lj, T& xi_out) {
  using N = Num<T>;
  const T dx = b.x1 - b.x0, dy = b.y1 - b.y0;
  const T s = dy / dx;
  const T yr = y - b.y0;
  const T t = b.d1 + b.d0 - 2 * s;
  const T qa = dy * (s - b.d0) + yr * t;
  const T qb = dy * b.d0 - yr * t;
  const T qc = -s * yr;
  const T xi = 2 * qc / (-qb - N::sqrt(qb * qb - 4 * qa * qc));
  const T om = 1 - xi;
  const T den = s + t * xi * om;
  x = xi * dx + b.x0;
  lj = -(N::log(N::abs(s * s * (b.d1 * xi * xi + 2 * s * xi * om + b.d0 * om * om))) - 2 * N::log(N::abs(den)));
  xi_out = xi;
}

// Evaluation of one (sample, coordinate) spline from its 3K-1 logits, which stay in the thread's private row of the
// shared-memory staging tile (row stride 3K-1 words is odd for every K: conflict-free).  exp() of each width/height
// logit is computed once and written back in place, so the bin search, the knots and the softmax backward all read
// the numerators from shared memory and nothing indexed dynamically lives in registers or local memory.
// On return row[0..2K) = exp(logit); row[2K..3K-1) still holds the derivative logits.
template <typename T, int KMAX, bool INV>
__device__ __forceinline__ void rqs_locate_row(T* __restrict__ row, int K, T B, T v, RqsBin<T>& b) {
  using N = Num<T>;
  T Sw = 0, Sh = 0;
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (k < K) {
      const T ew = N::exp(row[k]), eh = N::exp(row[K + k]);
      row[k] = ew; row[K + k] = eh;
      Sw = add_rn(Sw, ew); Sh = add_rn(Sh, eh);
    }
  b.Sw = Sw; b.Sh = Sh;
  const T* es = INV ? row + K : row;
  const T* eo = INV ? row : row + K;
  const T rSs = 1 / (INV ? Sh : Sw), rSo = 1 / (INV ? Sw : Sh);   // softmax by reciprocal multiply (one division per family)
  const T twoB = 2 * B;
  // knots of the searched family: sequential cumsum, knot = (2B)*cs - B without FMA contraction
  T cs = 0, prev = -B, prevc = 0, s0 = 0, s1 = 0, c0 = 0, c1 = 0;
  int bin = K + 1;
  if (!(prev < v)) bin = 0;
  else
    for (int k = 1; k <= K; ++k) {
      cs = add_rn(cs, mul_rn(es[k - 1], rSs));
      const T knot = add_rn(mul_rn(twoB, cs), -B);
      if (!(knot < v)) { bin = k; s0 = prev; s1 = knot; c0 = prevc; c1 = cs; break; }
      prev = knot; prevc = cs;
    }
  b.k = bin;
  if (bin < 1 || bin > K) return;
  T co = 0, o0 = -B, o1 = -B, oc0 = 0, oc1 = 0;
  for (int k = 1; k <= bin; ++k) {
    o0 = o1; oc0 = oc1;
    co = add_rn(co, mul_rn(eo[k - 1], rSo));
    o1 = add_rn(mul_rn(twoB, co), -B); oc1 = co;
  }
  if (!INV) { b.x0 = s0; b.x1 = s1; b.cx0 = c0; b.cx1 = c1; b.y0 = o0; b.y1 = o1; b.cy0 = oc0; b.cy1 = oc1; }
  else      { b.y0 = s0; b.y1 = s1; b.cy0 = c0; b.cy1 = c1; b.x0 = o0; b.x1 = o1; b.cx0 = oc0; b.cx1 = oc1; }
  b.d0 = (bin - 1 == 0) ? T(1) : N::log(N::exp(row[2 * K + bin - 2]) + 1);
  b.d1 = (bin == K) ? T(1) : N::log(N::exp(row[2 * K + bin - 1]) + 1);
}

// ---- bulk-copy (TMA 1-D) staging for the spline kernels -------------------------------------------------------
// A block's tile = blockDim.x consecutive (sample, coordinate) pairs = blockDim.x * (3K-1) contiguous logits: one
// cp.async.bulk per tile into a ring of shared-memory buffers (mbarrier complete_tx), issued RQS ring-depth - 1 tiles
// ahead so the HBM latency hides behind the arithmetic of the current tile; gradients leave the same way (bulk store).
namespace rq {
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void load_tile(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void store_tile(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void wait_stores_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > (1u << 28)) __trap();   // fail loudly instead of hanging the GPU
  }
}
constexpr int kMaxStages = 4;
}  // namespace rq

template <typename T, int KMAX, bool INV>
__device__ __forceinline__ void rqs_apply_row(T* __restrict__ row, int64_t r, int i, const T* __restrict__ Xin, const int* __restrict__ idx1,
                                              int c, int d, int K, T B, T* __restrict__ Xout, int32_t* __restrict__ bins,
                                              float& run_max, T& lj_out, int64_t& r_out, int& i_out) {
  const int64_t e = r * c + i;
  const int j = idx1[i];
  const T v = Xin[r * d + j];
  RqsBin<T> b;
  rqs_locate_row<T, KMAX, INV>(row, K, B, v, b);
  if (bins) bins[e] = b.k;
  T o = v, lj = 0, tmp;
  if (b.k >= 1 && b.k <= K) { if (!INV) rqs_eval_fwd(b, v, o, lj); else rqs_eval_inv(b, v, o, lj, tmp); }
  Xout[r * d + j] = o;
  run_max = fmaxf(run_max, fabsf((float)o));
  lj_out = lj; r_out = r; i_out = i;
}

// One thread per (sample, transformed coordinate); tiles of blockDim.x pairs stream through a `stages`-deep ring of
// shared-memory buffers (see rq:: above).  Passthrough columns are copied by a separate coalesced loop.
// Dynamic smem: stages * blockDim.x * (3K-1) * sizeof(T).
template <typename T, int KMAX, bool INV>
__global__ void rqs_apply_kernel(const T* __restrict__ Xin, const T* __restrict__ raw, const int* __restrict__ idx1,
                                 const int* __restrict__ pos, int c, int d, int K, T B, int64_t N, T* __restrict__ Xout,
                                 T* __restrict__ ld, int32_t* __restrict__ bins, float* __restrict__ amax_meta, int stages) {
  extern __shared__ __align__(128) unsigned char rqs_smem[];
  __shared__ uint64_t full_bar[rq::kMaxStages];
  T* tiles = reinterpret_cast<T*>(rqs_smem);
  const int P3 = 3 * K - 1;
  const int tile_elems = blockDim.x * P3;
  const uint32_t tile_bytes = (uint32_t)(tile_elems * sizeof(T));
  const int64_t total = N * c;
  const int64_t nfull = total / blockDim.x;
  float run_max = 0.f;
  if (threadIdx.x == 0) {
    for (int q = 0; q < stages; ++q) rq::bar_init(rq::s32(&full_bar[q]));
    rq::bar_fence_init();
  }
  __syncthreads();
  const int64_t n_my = blockIdx.x < nfull ? (nfull - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (threadIdx.x == 0)
    for (int q = 0; q < stages - 1 && q < n_my; ++q)
      rq::load_tile(rq::s32(tiles + (size_t)q * tile_elems), raw + (blockIdx.x + (int64_t)q * gridDim.x) * tile_elems, tile_bytes,
                    rq::s32(&full_bar[q]));
  // passthrough columns (overlaps the first tiles' flight)
  {
    const int64_t nd = N * d;
    const bool small = nd < ((int64_t)1 << 31);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nd; e += (int64_t)gridDim.x * blockDim.x) {
      const int j = small ? (int)((unsigned int)e % (unsigned int)d) : (int)(e % d);
      if (pos[j] < 0) { const T x = Xin[e]; Xout[e] = x; run_max = fmaxf(run_max, fabsf((float)x)); }
    }
  }
  const bool pow2 = (c & (c - 1)) == 0 && c <= 32 && (blockDim.x % c) == 0;
  const bool whole_rows = (blockDim.x % c) == 0;          // a tile is whole samples: (row, coordinate) of a thread is loop-invariant
  const int rows_per_tile = blockDim.x / c, t_row = threadIdx.x / c, t_col = threadIdx.x % c;
  for (int64_t it = 0; it < n_my; ++it) {
    const int buf = (int)(it % stages);
    const int64_t blk = blockIdx.x + it * gridDim.x;
    // refill the buffer tile it-1 used (every thread finished with it before the barrier below; the proxy fence orders
    // the in-place exp() writes before the bulk copy that overwrites them)
    rq::fence_async();
    __syncthreads();
    if (threadIdx.x == 0 && it + stages - 1 < n_my) {
      const int nb = (int)((it + stages - 1) % stages);
      rq::load_tile(rq::s32(tiles + (size_t)nb * tile_elems), raw + (blk + (int64_t)(stages - 1) * gridDim.x) * tile_elems, tile_bytes,
                    rq::s32(&full_bar[nb]));
    }
    rq::wait(rq::s32(&full_bar[buf]), (uint32_t)((it / stages) & 1));
    T lj_out = 0; int64_t r_out = 0; int i_out = -1;
    int64_t r; int i;
    if (whole_rows) { r = blk * rows_per_tile + t_row; i = t_col; }      // no 64-bit division on the hot path
    else { const int64_t e = blk * blockDim.x + threadIdx.x; r = e / c; i = (int)(e - r * c); }
    rqs_apply_row<T, KMAX, INV>(tiles + (size_t)buf * tile_elems + threadIdx.x * P3, r, i, Xin, idx1, c, d, K, B, Xout, bins, run_max,
                                lj_out, r_out, i_out);
    if (ld) {
      if (pow2) {          // the c coordinates of a sample sit in c adjacent lanes: segmented butterfly, one add per sample
        for (int o = c >> 1; o > 0; o >>= 1) lj_out += __shfl_xor_sync(0xffffffffu, lj_out, o);
        if (i_out == 0) atomicAdd(&ld[r_out], lj_out);
      } else if (lj_out != T(0)) {
        atomicAdd(&ld[r_out], lj_out);
      }
    }
  }
  // ragged tail (total % blockDim.x pairs): plain staging by the block whose turn it would be
  const int rem = (int)(total - nfull * blockDim.x);
  if (rem > 0 && blockIdx.x == (unsigned)(nfull % gridDim.x)) {
    __syncthreads();
    const int64_t e0 = nfull * blockDim.x;
    for (int i = threadIdx.x; i < rem * P3; i += blockDim.x) tiles[i] = raw[e0 * P3 + i];
    __syncthreads();
    if ((int)threadIdx.x < rem) {
      T lj_out = 0; int64_t r_out = 0; int i_out = -1;
      const int64_t e = e0 + threadIdx.x;
      rqs_apply_row<T, KMAX, INV>(tiles + threadIdx.x * P3, e / c, (int)(e % c), Xin, idx1, c, d, K, B, Xout, bins, run_max, lj_out, r_out, i_out);
      if (ld && lj_out != T(0)) atomicAdd(&ld[r_out], lj_out);
    }
  }
  if (amax_meta) amax_update(amax_meta, run_max);
}

// One row of the backward: turns the thread's private row of logits into its row of gradients in place.
template <typename T, int KMAX, bool INV>
__device__ __forceinline__ void rqs_bwd_row(T* __restrict__ row, int64_t r, int i, T* __restrict__ G, const T* __restrict__ Vsrc,
                                            const T* __restrict__ gld, const int* __restrict__ idx1, int c, int d, int K, T B,
                                            float& run_max, const T* __restrict__ gsrc, T* __restrict__ gkeep, bool store_gin) {
  // gsrc: where the incoming gradient of this coordinate is read (G[r, idx1[i]]; the side buffer in a redo pass, because the
  // speculative pass has already overwritten G); gkeep: optional copy of it for such a redo; d/d(spline input) always goes to
  // G[r, idx1[i]] unless store_gin is false (estimate pass: nothing is stored)
  using Nm = Num<T>;
  const int P3 = 3 * K - 1;
  const int j = idx1[i];
  const T v = Vsrc[r * d + j];
  const T go = *gsrc;
  if (gkeep) *gkeep = go;
  const T gl = gld ? gld[r] : T(1);
  RqsBin<T> b;
  rqs_locate_row<T, KMAX, INV>(row, K, B, v, b);
  if (b.k >= 1 && b.k <= K) {
    const T dx = b.x1 - b.x0, dy = b.y1 - b.y0;
    const T idx_ = 1 / dx;
    const T s = dy * idx_;
    T x, xi;
    if (!INV) { x = v; xi = (x - b.x0) * idx_; }
    else { T lj_; rqs_eval_inv(b, v, x, lj_, xi); }
    const T om = 1 - xi;
    const T tt = b.d1 + b.d0 - 2 * s;
    const T den = s + tt * xi * om;
    const T iden = 1 / den;
    const T num = s * xi * xi + b.d0 * xi * om;
    const T q = b.d1 * xi * xi + 2 * s * xi * om + b.d0 * om * om;
    const T iq = 1 / q;
    const T dq_dxi = 2 * b.d1 * xi + 2 * s * (om - xi) - 2 * b.d0 * om;
    T gy, glj, g_in = 0;
    if (!INV) { gy = go; glj = gl; }
    else {
      // L = go*x - gl*lj(x,p):  dL/dy = (go - gl*lj_x)/F_x =: a,  dL/dp = -a*F_p - gl*lj_p
      const T lj_xi = dq_dxi * iq - 2 * tt * (om - xi) * iden;
      const T Fx = s * s * q * iden * iden;
      const T a = (go - gl * (lj_xi * idx_)) / Fx;
      gy = -a; glj = -gl; g_in = a;
    }
    T g_y0 = gy, g_dy = gy * num * iden;
    const T g_num = gy * dy * iden;
    T g_den = -gy * dy * num * iden * iden - 2 * glj * iden;
    T g_s = 2 * glj / s;
    const T g_q = glj * iq;
    T g_d1 = g_q * xi * xi, g_d0 = g_q * om * om;
    g_s += g_q * 2 * xi * om;
    T g_xi = g_q * dq_dxi;
    g_s += g_num * xi * xi; g_d0 += g_num * xi * om; g_xi += g_num * (2 * s * xi + b.d0 * (om - xi));
    g_s += g_den * (1 - 2 * xi * om); g_d1 += g_den * xi * om; g_d0 += g_den * xi * om; g_xi += g_den * tt * (om - xi);
    T g_x0 = -g_xi * idx_, g_dx = -g_xi * xi * idx_;
    if (!INV) g_in = g_xi * idx_;
    g_dy += g_s * idx_; g_dx += -g_s * s * idx_;
    const T g_y1 = g_dy; g_y0 -= g_dy;
    const T g_x1 = g_dx; g_x0 -= g_dx;
    if (store_gin) G[r * d + j] = g_in;
    const T twoB = 2 * B;
    const T rSw = 1 / b.Sw, rSh = 1 / b.Sh;
    const int bk = b.k;
    {   // knots -> width logits: X_j = 2B c_j - B, c_j = sum_{k<=j} p_k, p = softmax  (row[k] holds exp(logit))
      const T a0 = (bk - 1 >= 1) ? g_x0 : T(0), a1 = g_x1;
      const T dotp = twoB * (a0 * b.cx0 + a1 * b.cx1);
      const T gp0 = (twoB * (a0 + a1) - dotp) * rSw, gp1 = (twoB * a1 - dotp) * rSw, gp2 = -dotp * rSw;
#pragma unroll
      for (int k = 1; k <= KMAX; ++k) if (k <= K) {
        const T gv = row[k - 1] * (k <= bk - 1 ? gp0 : (k <= bk ? gp1 : gp2));
        row[k - 1] = gv; run_max = fmaxf(run_max, fabsf((float)gv));
      }
    }
    {   // height logits
      const T a0 = (bk - 1 >= 1) ? g_y0 : T(0), a1 = g_y1;
      const T dotp = twoB * (a0 * b.cy0 + a1 * b.cy1);
      const T gp0 = (twoB * (a0 + a1) - dotp) * rSh, gp1 = (twoB * a1 - dotp) * rSh, gp2 = -dotp * rSh;
#pragma unroll
      for (int k = 1; k <= KMAX; ++k) if (k <= K) {
        const T gv = row[K + k - 1] * (k <= bk - 1 ? gp0 : (k <= bk ? gp1 : gp2));
        row[K + k - 1] = gv; run_max = fmaxf(run_max, fabsf((float)gv));
      }
    }
    T gd0 = 0, gd1 = 0;   // through softplus: d/dlogit log(exp(t)+1) = exp(t)/(exp(t)+1)
    if (bk - 1 >= 1) { const T ex = Nm::exp(row[2 * K + bk - 2]); gd0 = g_d0 * ex / (ex + 1); }
    if (bk <= K - 1) { const T ex = Nm::exp(row[2 * K + bk - 1]); gd1 = g_d1 * ex / (ex + 1); }
    run_max = fmaxf(run_max, fmaxf(fabsf((float)gd0), fabsf((float)gd1)));
#pragma unroll
    for (int k = 1; k < KMAX; ++k) if (k < K)
      row[2 * K + k - 1] = (k == bk - 1) ? gd0 : ((k == bk) ? gd1 : T(0));
  } else {   // identity tails: dy/dx = 1, no parameter gradient; G unchanged
    if (store_gin && gsrc != G + (r * d + j)) G[r * d + j] = go;
#pragma unroll
    for (int k = 0; k < 3 * KMAX - 1; ++k) if (k < P3) row[k] = 0;
  }
}

// Optional second output form of the spline backward (Float32 tcgen05 path): the gradient w.r.t. the conditioner output goes
// straight into the fp16 hi / lo planes the dgrad / weight-gradient GEMMs read, instead of an fp32 matrix that a separate
// pass would re-read and split.  The planes need ONE power-of-two scale for the whole tensor, i.e. max |gradient| before the
// first store.  It is predicted: a dry pass over every tile_step-th tile records the maximum of that sample (est), the
// full pass scales with 2^8 of headroom over it, records the exact maximum and keeps a copy of the incoming gradients it
// overwrites, and a third launch returns at once -- or, if the exact maximum would have overflowed the predicted scale,
// redoes the tile loop with the exact scale from that copy.  Every decision is a pure function of the inputs, so results stay reproducible;
// a scale up to 2^8 below the ideal one costs nothing visible (error floor 2^-32 of the maximum instead of 2^-40).
template <typename T> struct RqsPlanesOut {
  __half* hi;             // hi plane [rows_pad, ld]; lo plane at + plane_elems
  int64_t plane_elems;
  int ld;
  float* meta;            // {scale, bits of the exact max |gradient|} of the planes
  float* est;             // est[1]: bits of the max over the sampled tiles
  T* gside;               // [N, c] incoming gradient of every transformed coordinate, kept by the speculative pass for a redo
  int mode;               // 0: fp32 graw (no planes); 1: estimate pass; 2: speculative pass; 3: nothing, or redo with the exact scale
  int tile_step;          // mode 1: every tile_step-th tile
  float headroom;         // predicted bound = headroom * sampled maximum (kRqsSpecHeadroom; tests force a redo with a tiny value)
};
constexpr float kRqsSpecHeadroom = 256.f;
__device__ __forceinline__ float rqs_pow2_scale(float bound) {     // same rule as the tensor path: bound * s <= 2^14
  if (!(bound > 0.f) || !(bound < 3.0e38f)) return 1.f;
  int sh = 14 - (ilogbf(bound) + 1);
  sh = sh > 100 ? 100 : (sh < -100 ? -100 : sh);
  return ldexpf(1.f, sh);
}
__device__ __forceinline__ void rqs_split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// quad form of rqs_store_planes: Float32 rows of a multiple of four columns, and the block covers whole rows of quads
__device__ __forceinline__ bool rqs_planes_quads(int ncol, int ld) {
  const int quads = ld >> 2;
  return (ncol & 3) == 0 && quads <= (int)blockDim.x && (blockDim.x % quads) == 0;
}
// rows x ncol fp32 gradients of whole samples r0 .. r0 + rows - 1 (shared memory) -> scaled hi / lo planes (padding columns
// zero), and the column sums (= bias gradient of the conditioner's last Dense): thread t owns column pairs t, t + blockDim
template <typename T>
__device__ __forceinline__ void rqs_store_planes(const T* __restrict__ tile, int rows, int64_t r0, int ncol, float s,
                                                 const RqsPlanesOut<T>& po, T (&csum)[4], bool want_sums) {
  const int half_ld = po.ld >> 1;
  if (rqs_planes_quads(ncol, po.ld)) {
    // four adjacent columns per thread (one 16-byte shared load, one 8-byte store per plane), blockDim / (ld / 4) rows per pass
    const int quads = po.ld >> 2, qd = threadIdx.x % quads, rstep = blockDim.x / quads;
    const int col = 4 * qd;
    const bool live = col < ncol;
    T a[4] = {0, 0, 0, 0};
    for (int rr = threadIdx.x / quads; rr < rows; rr += rstep) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) v = *reinterpret_cast<const float4*>(tile + rr * ncol + col);
      a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
      uint2 hi, lo;
      rqs_split_pair(v.x * s, v.y * s, hi.x, lo.x);
      rqs_split_pair(v.z * s, v.w * s, hi.y, lo.y);
      const int64_t o = (r0 + rr) * po.ld + col;
      *reinterpret_cast<uint2*>(po.hi + o) = hi;
      *reinterpret_cast<uint2*>(po.hi + po.plane_elems + o) = lo;
    }
    if (want_sums) { csum[0] += a[0]; csum[1] += a[1]; csum[2] += a[2]; csum[3] += a[3]; }
    return;
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int cp = threadIdx.x + q * blockDim.x;
    if (cp < half_ld) {
      const int col = 2 * cp;
      const bool live = col < ncol;          // ncol is even: a pair is all live or all padding
      T a0 = 0, a1 = 0;
      uint32_t* ph = reinterpret_cast<uint32_t*>(po.hi + r0 * po.ld + col);
      uint32_t* pl = reinterpret_cast<uint32_t*>(po.hi + po.plane_elems + r0 * po.ld + col);
      for (int rr = 0; rr < rows; ++rr) {
        T v0 = 0, v1 = 0;
        if (live) { v0 = tile[rr * ncol + col]; v1 = tile[rr * ncol + col + 1]; }
        a0 += v0; a1 += v1;
        uint32_t hi, lo;
        rqs_split_pair((float)v0 * s, (float)v1 * s, hi, lo);
        ph[(int64_t)rr * half_ld] = hi;
        pl[(int64_t)rr * half_ld] = lo;
      }
      if (want_sums) { csum[2 * q] += a0; csum[2 * q + 1] += a1; }
    }
  }
}

// Backward of the spline coupling arithmetic: G (in place on idx1 columns) and graw = d/dtheta_raw.
// Vsrc = the spline input (Xin): x1 for forward, y1 for inverse.  Tiles stream in through the bulk-copy ring, every
// thread rewrites its private row in place, and the tile streams back out with one bulk store.
// colsum (optional, needs blockDim % c == 0 and c*(3K-1) <= 4*blockDim): sum over samples of every graw column = the
// bias gradient of the conditioner's last Dense, accumulated per thread across the block's tiles and flushed once.
template <typename T, int KMAX, bool INV, int MODE>
__global__ void __launch_bounds__(128, (KMAX <= 16 && sizeof(T) == 4) ? 7 : 4)   // Float32, K <= 16: seven 128-thread blocks per SM fit the ring
rqs_bwd_kernel(T* __restrict__ G, const T* __restrict__ Vsrc, const T* __restrict__ raw,
                               const T* __restrict__ gld, const int* __restrict__ idx1, int c, int d, int K, T B,
                               int64_t N, T* __restrict__ graw, float* __restrict__ amax_meta, double* __restrict__ colsum,
                               int stages, RqsPlanesOut<T> po) {
  extern __shared__ __align__(128) unsigned char rqs_smem[];
  __shared__ uint64_t full_bar[rq::kMaxStages];
  T* tiles = reinterpret_cast<T*>(rqs_smem);
  const int P3 = 3 * K - 1;
  const int ncol = c * P3;
  const int tile_elems = blockDim.x * P3;
  const uint32_t tile_bytes = (uint32_t)(tile_elems * sizeof(T));
  const int64_t total = N * c;
  const int64_t nfull = total / blockDim.x;
  constexpr int mode = MODE;      // a template parameter: the fp32 form keeps its register count (occupancy)
  // planes modes: the scale of this pass, and (mode 3) whether the speculative pass has to be redone at all
  float s_planes = 1.f;
  if (mode == 2) s_planes = rqs_pow2_scale(__uint_as_float(reinterpret_cast<const unsigned int*>(po.est)[1]) * po.headroom);
  if (mode == 3) {
    const float s_spec = rqs_pow2_scale(__uint_as_float(reinterpret_cast<const unsigned int*>(po.est)[1]) * po.headroom);
    s_planes = rqs_pow2_scale(__uint_as_float(reinterpret_cast<const unsigned int*>(po.meta)[1]));
    const unsigned int est_bits = reinterpret_cast<const unsigned int*>(po.est)[1], exact_bits = reinterpret_cast<const unsigned int*>(po.meta)[1];
    // redo when the exact maximum overflows the predicted scale, or when the sample saw only zeros (identity tails) and
    // the prediction therefore says nothing about the magnitude
    const bool redo = (s_planes < s_spec) || (est_bits == 0u && exact_bits != 0u);
    if (!redo) return;       // the predicted scale held: nothing to do
  }
  if ((mode == 2 || mode == 3) && blockIdx.x == 0 && threadIdx.x == 0) po.meta[0] = s_planes;
  const bool planes = mode >= 2;
  const bool want_sums = colsum && mode != 3 && mode != 1;     // a redo repeats the planes, not the (scale-free) column sums
  // mode 1 walks every tile_step-th tile and stores nothing
  const int64_t step = mode == 1 ? (int64_t)(po.tile_step > 0 ? po.tile_step : 1) : 1;
  const int64_t ntl = (nfull + step - 1) / step;
  float run_max = 0.f;
  T csum[4] = {0, 0, 0, 0};
  if (threadIdx.x == 0) {
    for (int q = 0; q < stages; ++q) rq::bar_init(rq::s32(&full_bar[q]));
    rq::bar_fence_init();
  }
  __syncthreads();
  const int64_t n_my = blockIdx.x < ntl ? (ntl - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (threadIdx.x == 0)
    for (int q = 0; q < stages - 1 && q < n_my; ++q)
      rq::load_tile(rq::s32(tiles + (size_t)q * tile_elems), raw + (blockIdx.x + (int64_t)q * gridDim.x) * step * tile_elems, tile_bytes,
                    rq::s32(&full_bar[q]));
  const bool whole_rows = (blockDim.x % c) == 0;
  const int rows_per_tile = blockDim.x / c, t_row = threadIdx.x / c, t_col = threadIdx.x % c;
  for (int64_t it = 0; it < n_my; ++it) {
    const int buf = (int)(it % stages);
    const int64_t blk = (blockIdx.x + it * gridDim.x) * step;
    T* tile = tiles + (size_t)buf * tile_elems;
    rq::wait(rq::s32(&full_bar[buf]), (uint32_t)((it / stages) & 1));
    {
      int64_t r; int i;
      if (whole_rows) { r = blk * rows_per_tile + t_row; i = t_col; }    // no 64-bit division on the hot path
      else { const int64_t e = blk * blockDim.x + threadIdx.x; r = e / c; i = (int)(e - r * c); }
      // speculative pass: keeps the incoming gradient in the side buffer; a redo reads it from there (G is already overwritten)
      const T* gsrc = mode == 3 ? po.gside + (r * c + i) : G + (r * d + idx1[i]);
      T* gkeep = mode == 2 ? po.gside + (r * c + i) : nullptr;
      rqs_bwd_row<T, KMAX, INV>(tile + threadIdx.x * P3, r, i, G, Vsrc, gld, idx1, c, d, K, B, run_max, gsrc, gkeep, mode != 1);
    }
    rq::fence_async();       // generic-proxy accesses of the rows -> ordered before the bulk store / the refill of this buffer
    __syncthreads();
    if (threadIdx.x == 0) {
      if (mode == 0) rq::store_tile(graw + blk * tile_elems, rq::s32(tile), tile_bytes);
      if (it + stages - 1 < n_my) {
        // the buffer of tile it-1 is next in the ring: its store must have finished reading shared memory (every
        // thread's reads of it -- column sums, plane stores -- completed before the barrier above)
        if (mode == 0) rq::wait_stores_read<1>();
        const int nb = (int)((it + stages - 1) % stages);
        rq::load_tile(rq::s32(tiles + (size_t)nb * tile_elems), raw + (blk + (int64_t)(stages - 1) * gridDim.x * step) * tile_elems, tile_bytes,
                      rq::s32(&full_bar[nb]));
      }
    }
    if (planes) {            // the tile is whole samples (the launcher checks): rows of ncol gradients -> planes (+ column sums)
      rqs_store_planes<T>(tile, tile_elems / ncol, blk * rows_per_tile, ncol, s_planes, po, csum, want_sums);
    } else if (want_sums) {   // the tile is whole samples: blockDim / c rows of ncol columns
      const int rows = tile_elems / ncol;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = threadIdx.x + q * blockDim.x;
        if (col < ncol) {
          T sacc = 0;
          for (int rr = 0; rr < rows; ++rr) sacc += tile[rr * ncol + col];
          csum[q] += sacc;
        }
      }
    }
  }
  if (threadIdx.x == 0 && mode == 0) rq::wait_stores_read<0>();
  // ragged tail (total % blockDim.x pairs): plain staging by the block whose turn it would be (not part of the estimate)
  const int rem = (int)(total - nfull * blockDim.x);
  if (rem > 0 && mode != 1 && blockIdx.x == (unsigned)(nfull % gridDim.x)) {
    __syncthreads();
    const int64_t e0 = nfull * blockDim.x;
    for (int i = threadIdx.x; i < rem * P3; i += blockDim.x) tiles[i] = raw[e0 * P3 + i];
    __syncthreads();
    if ((int)threadIdx.x < rem) {
      const int64_t e = e0 + threadIdx.x;
      const int64_t r = e / c; const int i = (int)(e % c);
      const T* gsrc = mode == 3 ? po.gside + e : G + (r * d + idx1[i]);
      T* gkeep = mode == 2 ? po.gside + e : nullptr;
      rqs_bwd_row<T, KMAX, INV>(tiles + threadIdx.x * P3, r, i, G, Vsrc, gld, idx1, c, d, K, B, run_max, gsrc, gkeep, true);
    }
    __syncthreads();
    if (planes) {
      rqs_store_planes<T>(tiles, rem * P3 / ncol, e0 / c, ncol, s_planes, po, csum, want_sums);
    } else {
      for (int i = threadIdx.x; i < rem * P3; i += blockDim.x) graw[e0 * P3 + i] = tiles[i];
      if (want_sums) {
        const int rows = rem * P3 / ncol;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = threadIdx.x + q * blockDim.x;
          if (col < ncol) for (int rr = 0; rr < rows; ++rr) csum[q] += tiles[rr * ncol + col];
        }
      }
    }
  }
  if (mode == 1) { amax_update(po.est, run_max); return; }
  if (amax_meta && mode != 3) amax_update(amax_meta, run_max);
  if (want_sums) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      // fp32 form: thread t owns columns t + q * blockDim; planes form: column pairs t, t + blockDim (rqs_store_planes)
      const int col = !planes ? threadIdx.x + q * blockDim.x
                      : (rqs_planes_quads(ncol, po.ld) ? 4 * (threadIdx.x % (po.ld >> 2)) + q
                                                       : 2 * (threadIdx.x + (q >> 1) * blockDim.x) + (q & 1));
      if (col < ncol) atomicAdd(&colsum[col], (double)csum[q]);
    }
  }
}

template <typename T>
__global__ void rqs_bin_search_kernel(const T* __restrict__ knots, const T* __restrict__ v, int64_t M, int K, int32_t* __restrict__ bins) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M) return;
  const T* kn = knots + e * (K + 1);
  int k = 0;
  for (int i = 0; i <= K; ++i) k += (kn[i] < v[e]) ? 1 : 0;
  bins[e] = k;
}

// ---------------------------------------------------------------------------------------------
// objective heads
// ---------------------------------------------------------------------------------------------
// ELBO head, coalesced: a CTA stages 128 rows of X0 and then of Y through shared memory (pitch d+1: conflict-free
// row-per-thread access), evaluates the target in place (z -> dlogp/dz) and streams G back out.
// Dynamic smem: blockDim.x * (d + 1) * sizeof(T).
template <typename T>
__global__ void elbo_head_tiled_kernel(const T* __restrict__ Y, const T* __restrict__ X0, const T* __restrict__ ld,
                                       TargetParams<T> tp, const T* __restrict__ base, T base_c0, int d, int64_t N,
                                       T* __restrict__ G, T* __restrict__ terms, double* __restrict__ sum_out,
                                       const T* __restrict__ lq0 = nullptr) {
  extern __shared__ __align__(16) unsigned char head_smem[];
  T* sm = reinterpret_cast<T*>(head_smem);
  const int nthr = blockDim.x, tid = threadIdx.x, pitch = d + 1;
  const int64_t r0 = (int64_t)blockIdx.x * nthr;
  const int nrows = (int)((N - r0) < nthr ? (N - r0) : nthr);
  const int cnt = nrows * d;
  T q = 0;
  if (!lq0) {          // lq0: per-sample log q0(x0) computed elsewhere (full-covariance base)
    for (int i = tid; i < cnt; i += nthr) sm[(i / d) * pitch + (i % d)] = X0[r0 * d + i];
    __syncthreads();
    if (tid < nrows)
      for (int k = 0; k < d; ++k) {
        const T x = sm[tid * pitch + k];
        const T u = base ? (x - base[k]) / base[d + k] : x;
        q += u * u;
      }
    __syncthreads();
  }
  for (int i = tid; i < cnt; i += nthr) sm[(i / d) * pitch + (i % d)] = Y[r0 * d + i];
  __syncthreads();
  double term = 0;
  if (tid < nrows) {
    const T lp = target_logp_score<T, 0>(tp, sm + tid * pitch, sm + tid * pitch);
    const T t = lp - (lq0 ? lq0[r0 + tid] : (base_c0 - q / 2)) + ld[r0 + tid];
    if (terms) terms[r0 + tid] = t;
    term = (double)t;
  }
  __syncthreads();
  for (int i = tid; i < cnt; i += nthr) G[r0 * d + i] = sm[(i / d) * pitch + (i % d)];
  term = warp_sum(term);
  __shared__ double sh[32];
  if ((tid & 31) == 0) sh[tid >> 5] = term;
  __syncthreads();
  if (tid == 0) {
    double s = 0;
    for (int w = 0; w < (nthr + 31) / 32; ++w) s += sh[w];
    atomicAdd(sum_out, s);
  }
}

// ELBO head (reference src/objectives/elbo.jl:65-70): term = logp(y) - logq0(x0) + logdet; G = dlogp/dy.
template <typename T>
__global__ void elbo_head_kernel(const T* __restrict__ Y, const T* __restrict__ X0, const T* __restrict__ ld,
                                 TargetParams<T> tp, const T* __restrict__ base, T base_c0, int d, int64_t N,
                                 T* __restrict__ G, T* __restrict__ terms, double* __restrict__ sum_out,
                                 const T* __restrict__ lq0 = nullptr) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double term = 0;
  if (r < N) {
    const T lp = target_logp_score<T, 0>(tp, Y + r * d, G + r * d);
    T q = 0;
    if (!lq0)
      for (int k = 0; k < d; ++k) {
        const T u = base ? (X0[r * d + k] - base[k]) / base[d + k] : X0[r * d + k];
        q += u * u;
      }
    const T t = lp - (lq0 ? lq0[r] : (base_c0 - q / 2)) + ld[r];
    if (terms) terms[r] = t;
    term = (double)t;
  }
  term = warp_sum(term);
  __shared__ double sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = term;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) s += sh[w];
    atomicAdd(sum_out, s);
  }
}

// log-likelihood head (reference src/objectives/loglikelihood.jl:26-33 via Bijectors logpdf):
// term = logq0(x0) + logdet_inv; G = dlogq0/dx0.
template <typename T>
__global__ void loglik_head_kernel(const T* __restrict__ X0, const T* __restrict__ ld, const T* __restrict__ base,
                                   T base_c0, int d, int64_t N, T* __restrict__ G, T* __restrict__ terms,
                                   double* __restrict__ sum_out, const T* __restrict__ lq0 = nullptr) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double term = 0;
  if (r < N) {
    T q = 0;
    if (!lq0)      // lq0 (+ G already written): full-covariance base handled by base_dense_kernel
    for (int k = 0; k < d; ++k) {
      const T is = base ? 1 / base[d + k] : T(1);
      const T u = base ? (X0[r * d + k] - base[k]) * is : X0[r * d + k];
      q += u * u;
      if (G) G[r * d + k] = -u * is;
    }
    const T t = (lq0 ? lq0[r] : (base_c0 - q / 2)) + ld[r];
    if (terms) terms[r] = t;
    term = (double)t;
  }
  term = warp_sum(term);
  __shared__ double sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = term;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) s += sh[w];
    atomicAdd(sum_out, s);
  }
}

}  // namespace nf
