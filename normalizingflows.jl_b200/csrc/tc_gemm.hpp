// tcgen05 / TMEM GEMM path for the coupling-MLP contractions of Float32 flows (kernel K3).
#pragma once
#include "flow.hpp"

namespace nf {

// bytes of a split-bf16 activation buffer (hi and lo planes, rows padded to 128, features to 64)
size_t tc_act_bytes(int64_t n, int width);
// bytes of the per-flow prepared weight planes
size_t tc_weight_bytes(const Flow& f);
// theta -> transposed / split bf16 weight planes (once per call, before any tc_mlp_*)
int tc_prepare_weights(Flow& f, const float* theta_dev);
// forget the per-tensor scale slots of the previous sample chunk (buffers are about to be reused)
int tc_begin_chunk(Flow& f);
// x2 = X[:, idx2] -> split planes
// amax_src: optional metadata slot whose [1] bounds max |X| (skips the absmax pass)
int tc_gather_split(Flow& f, const float* X, int d, const int* d_idx, int n_idx, int64_t n, void* act0, const float* amax_src);
// fresh zeroed {scale, amax} slot; exact max |X| over `count` floats
float* tc_alloc_meta(Flow& f);
int tc_absmax(Flow& f, const float* X, int64_t count, float* meta);
int tc_mlp_forward(Flow& f, const LayerDesc& Ld, int m, int64_t n, void* act0, std::vector<void*>& acts);
// g_last: fp32 [n, out] gradient w.r.t. the last Dense's pre-activation; scratch0/1: activation-sized buffers.
// g_last == nullptr: the producer has already written that gradient as split planes into scratch0 (tc_planes_out) together
// with its bias-gradient column sums.
int tc_mlp_backward(Flow& f, const LayerDesc& Ld, int m, int64_t n, void* act0, std::vector<void*>& acts, float* g_last,
                    const float* g_last_amax, void* scratch0, void* scratch1, float* G, double* gsum, bool last_bias_done = false);
// where a producer kernel writes split planes of an [n, width] tensor into `buf` itself: plane pointers, row stride and a fresh
// {scale, amax} slot the producer fills (scale = the power of two it multiplied with, amax = exact max |value|)
struct TcPlanesOut { void* hi; int64_t plane_elems; int ld; float* meta; };
int tc_planes_out(Flow& f, void* buf, int64_t n, int width, TcPlanesOut* out);
// fused AffineCoupling forward (both conditioners + coupling arithmetic in one launch); writes the same stash
// (x2 / hidden planes, sign bits, s) the layer-by-layer path writes, so tc_mlp_backward consumes it unchanged
bool tc_fused_affine_ok(const Flow& f, const LayerDesc& Ld);
int tc_affine_forward_fused(Flow& f, const LayerDesc& Ld, int64_t n, const float* Xin, float* Xout, float* ld, void* act0,
                            std::vector<std::vector<void*>>& acts, const float* x_meta, float* y_meta, bool inv, bool stash = true);
void tc_release(Flow& f);
int tc_gemm_selftest(int64_t n, int K, int N, const float* X_host, const float* Wt_host, const float* b_host, int terms,
                     float* Y_host);
// host-side item schedule of the two-team fused coupling kernel (fused_coupling.cuh: fused_build_schedule); no device needed
int tc_fused_schedule(int nch, int slab, int n_hoist, int delay, unsigned char* out, int cap);

}  // namespace nf
