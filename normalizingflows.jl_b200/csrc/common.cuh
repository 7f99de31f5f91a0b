// Shared host/device helpers for libnfcuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cmath>
#include <string>
#include <vector>
#include "../../include/nfcuda.h"

namespace nf {

// ---------------------------------------------------------------------------------------------
// error plumbing: nothing throws across the C ABI; every internal function returns nf_status.
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern thread_local int64_t g_launch_count;
extern int g_opt_fused_variant;    // 1 (default): 128-column MMAs (fused_coupling_w128.cuh); 0: the first, 64-column version
extern int g_opt_fused_coupling;   // nf_set_option("fused_coupling", .); initialised from NFCUDA_FUSED

#define NF_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      nf::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,    \
                    cudaGetErrorString(_e));                                                   \
      return (_e == cudaErrorMemoryAllocation) ? NF_ERR_OOM : NF_ERR_CUDA;                     \
    }                                                                                          \
  } while (0)

#define NF_TRY(expr)                 \
  do {                               \
    int _s = (expr);                 \
    if (_s != NF_OK) return _s;      \
  } while (0)

#define NF_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      nf::set_error(__VA_ARGS__);    \
      return NF_ERR_INVALID;         \
    }                                \
  } while (0)

#define NF_LAUNCH_CHECK()            \
  do {                               \
    ++nf::g_launch_count;            \
    NF_CUDA(cudaGetLastError());     \
  } while (0)

constexpr int kNumSMs = 148;  // B200

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---------------------------------------------------------------------------------------------
// device math in the flow's element type
// ---------------------------------------------------------------------------------------------
template <typename T> struct Num;
template <> struct Num<float> {
  static __device__ __forceinline__ float exp(float x) { return expf(x); }
  static __device__ __forceinline__ float log(float x) { return logf(x); }
  static __device__ __forceinline__ float log1p(float x) { return log1pf(x); }
  static __device__ __forceinline__ float tanh(float x) { return tanhf(x); }
  static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float abs(float x) { return fabsf(x); }
  static __device__ __forceinline__ float atan2(float y, float x) { return atan2f(y, x); }
  static __device__ __forceinline__ void sincos(float x, float* s, float* c) { sincosf(x, s, c); }
  static __device__ __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
};
template <> struct Num<double> {
  static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
  static __device__ __forceinline__ double log(double x) { return ::log(x); }
  static __device__ __forceinline__ double log1p(double x) { return ::log1p(x); }
  static __device__ __forceinline__ double tanh(double x) { return ::tanh(x); }
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ double abs(double x) { return ::fabs(x); }
  static __device__ __forceinline__ double atan2(double y, double x) { return ::atan2(y, x); }
  static __device__ __forceinline__ void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
  static __device__ __forceinline__ double max(double a, double b) { return ::fmax(a, b); }
};

// LogExpFunctions.log1pexp (stable softplus) and its derivative (logistic).
template <typename T> __host__ __device__ inline T softplus_stable(T x) {
  T ax = x < 0 ? -x : x;
  return (x > 0 ? x : T(0)) + (T)log1p(exp((double)-ax));
}
template <typename T> __host__ __device__ inline T sigmoid_stable(T x) {
  if (x >= 0) return T(1) / (T(1) + (T)exp((double)-x));
  T e = (T)exp((double)x);
  return e / (T(1) + e);
}

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Philox4x32-10 counter RNG (Salmon et al. 2011) -> standard normals via Box-Muller.  Replaces the
// CUDA.jl `randn!` of reference ext/NormalizingFlowsCUDAExt.jl:44 for device-generated base draws.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// both members of Box-Muller pair `pair` (elements 2 pair and 2 pair + 1 of the draw matrix): one Philox call, one log / sqrt /
// sincospi -- bit-identical to two philox_randn calls
template <typename T> __device__ __forceinline__ void philox_randn_pair(uint64_t seed, uint64_t pair, T& even, T& odd) {
  uint32_t r[4];
  philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), 0x6e66u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const double u1 = ((double)r[0] + 1.0) * (1.0 / 4294967296.0);
  const double u2 = (double)r[1] * (1.0 / 4294967296.0);
  const double rad = ::sqrt(-2.0 * ::log(u1));
  double s, c;
  ::sincospi(2.0 * u2, &s, &c);
  even = (T)(rad * c); odd = (T)(rad * s);
}
// element e of the N x dim draw matrix for (seed): deterministic, independent of launch geometry
template <typename T> __device__ __forceinline__ T philox_randn(uint64_t seed, uint64_t elem) {
  uint32_t r[4];
  const uint64_t pair = elem >> 1;
  philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), 0x6e66u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  // 2 uniforms in (0,1] / [0,1) from 2x32 bits each side -> one Box-Muller pair
  const double u1 = ((double)r[0] + 1.0) * (1.0 / 4294967296.0);
  const double u2 = (double)r[1] * (1.0 / 4294967296.0);
  const double rad = ::sqrt(-2.0 * ::log(u1));
  double s, c;
  ::sincospi(2.0 * u2, &s, &c);
  return (T)((elem & 1) ? rad * s : rad * c);
}

}  // namespace nf
