// Full-covariance base distribution q0 = MvNormal(mu, L L^T)  (the thing reference
// ext/NormalizingFlowsCUDAExt.jl:43-47 does on the GPU: randn -> unwhiten with the Cholesky factor -> + mu, and
// Distributions' logpdf(MvNormal) = -(d log 2pi + logdet Sigma)/2 - |L^-1 (x - mu)|^2 / 2).
//
// One thread per sample; a block stages its 128 rows (pitch d + 1: conflict free) and the factor in shared memory.
//   mode 0  X holds standard normal draws eps: lq = c0 - |eps|^2/2, then X <- mu + L eps          (sampling)
//   mode 1  X holds x0 (read only):            z = L^-1 (x0 - mu), lq = c0 - |z|^2/2, G <- -L^-T z  (log-density and its gradient)
// c0 = -d/2 log 2pi - sum_i log L_ii.  L is row major, lower triangular: L[i * d + k], k <= i.
#pragma once
#include "flow.hpp"

namespace nf {

// K7: base draws x = mu + sigma .* randn  (reference ext/NormalizingFlowsCUDAExt.jl:43-48)
// seed_iter (optional, device): added to the seed -- the iteration counter of a CUDA-graph-replayed training loop, so that one
// captured launch draws a fresh batch at every replay.
// One thread per Box-Muller PAIR (two consecutive elements of the global draw matrix share one Philox call and one set of
// transcendentals): half the work of one thread per element, same values.
template <typename T>
__global__ void base_sample_kernel(T* __restrict__ Z, const T* __restrict__ base, int d, int64_t N, uint64_t seed, int64_t row0,
                                   const int64_t* __restrict__ seed_iter = nullptr) {
  const int64_t first = row0 * d, count = N * d;            // this call owns global elements [first, first + count)
  const int64_t pair = (first >> 1) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * pair >= first + count) return;
  if (seed_iter) seed += (uint64_t)*seed_iter;
  T ze, zo;
  philox_randn_pair<T>(seed, (uint64_t)pair, ze, zo);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int64_t e = 2 * pair + u - first;
    if (e < 0 || e >= count) continue;
    T z = u ? zo : ze;
    if (base) { const int k = (int)(e % d); z = z * base[d + k] + base[k]; }
    Z[e] = z;
  }
}
// grid for base_sample_kernel (256 threads per block)
inline unsigned base_sample_grid(int64_t N, int d, int64_t row0) {
  const int64_t first = row0 * d, last = first + N * d;    // pairs first/2 .. (last - 1)/2
  const int64_t npairs = N * d > 0 ? ((last - 1) >> 1) - (first >> 1) + 1 : 0;
  return (unsigned)((npairs + 255) / 256);
}


template <typename T>
__global__ void __launch_bounds__(128) base_dense_kernel(T* __restrict__ X, const T* __restrict__ Lmu, int d, int64_t N, int mode, T c0,
                                                         T* __restrict__ lq, T* __restrict__ G) {
  extern __shared__ __align__(16) unsigned char bd_smem[];
  T* sL = reinterpret_cast<T*>(bd_smem);        // d*d
  T* smu = sL + d * d;                          // d
  T* rows = smu + d;                            // 128 x (d+1)
  const int tid = threadIdx.x, nthr = blockDim.x, pitch = d + 1;
  const int64_t r0 = (int64_t)blockIdx.x * nthr;
  const int nrows = (int)((N - r0) < nthr ? (N - r0) : nthr);
  for (int i = tid; i < d * d + d; i += nthr) sL[i] = Lmu[i];
  const int cnt = nrows * d;
  for (int i = tid; i < cnt; i += nthr) rows[(i / d) * pitch + (i % d)] = X[r0 * d + i];
  __syncthreads();
  T* v = rows + tid * pitch;
  if (tid < nrows) {
    if (mode == 0) {
      T q = 0;
      for (int k = 0; k < d; ++k) q += v[k] * v[k];
      if (lq) lq[r0 + tid] = c0 - q / 2;
      for (int i = d - 1; i >= 0; --i) {        // in place: row i only needs eps_k, k <= i, which are still untouched
        T s = smu[i];
        for (int k = 0; k <= i; ++k) s += sL[i * d + k] * v[k];
        v[i] = s;
      }
    } else {
      T q = 0;
      for (int i = 0; i < d; ++i) {             // forward substitution
        T s = v[i] - smu[i];
        for (int k = 0; k < i; ++k) s -= sL[i * d + k] * v[k];
        v[i] = s / sL[i * d + i];
        q += v[i] * v[i];
      }
      if (lq) lq[r0 + tid] = c0 - q / 2;
      if (G)
        for (int i = d - 1; i >= 0; --i) {      // back substitution with L^T
          T s = v[i];
          for (int k = i + 1; k < d; ++k) s -= sL[k * d + i] * v[k];
          v[i] = s / sL[i * d + i];
        }
    }
  }
  __syncthreads();
  if (mode == 0) { for (int i = tid; i < cnt; i += nthr) X[r0 * d + i] = rows[(i / d) * pitch + (i % d)]; }
  else if (G) { for (int i = tid; i < cnt; i += nthr) G[r0 * d + i] = -rows[(i / d) * pitch + (i % d)]; }
}

template <typename T>
inline int base_dense_launch(Flow& f, T* X, int64_t N, int mode, T* lq, T* G) {
  const int d = f.dim;
  const size_t smem = ((size_t)d * d + d + (size_t)128 * (d + 1)) * sizeof(T);
  auto kern = base_dense_kernel<T>;
  if (smem > 200 * 1024) { set_error("full-covariance base: dim %d needs %zu B of shared memory (limit 200 KiB)", d, smem); return NF_ERR_UNSUPPORTED; }
  if (smem > 48 * 1024) NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)ceil_div(N, 128), 128, smem, f.stream>>>(X, (const T*)f.d_base_L, d, N, mode, (T)f.base_c0, lq, G);
  NF_LAUNCH_CHECK();
  return NF_OK;
}

}  // namespace nf
