// Layered execution path for flows that contain coupling layers (RealNVP AffineCoupling,
// NeuralSplineCoupling): state X[N x d] streams through per-layer kernels, activations are stashed
// for the backward sweep.  See general.cu.
#pragma once
#include "flow.hpp"

namespace nf {

enum : int { OP_ELBO = 0, OP_FORWARD = 1, OP_INVERSE = 2, OP_LOGLIK = 3, OP_FORWARD_STASH = 4 };

struct GeneralJob {
  int op = OP_ELBO;
  const Target* tgt = nullptr;
  const void* theta_dev = nullptr;
  const void* in_dev = nullptr;
  int64_t N = 0;
  uint64_t seed = 0;
  const int64_t* seed_iter_dev = nullptr;   // device iteration counter added to the seed (graph-replayed training loop)
  bool want_grad = false;
  void* y_out = nullptr;
  void* ld_out = nullptr;
  void* terms_out = nullptr;
  int32_t* bins_out = nullptr;
};

// Grow the workspace so that a chunk of the batch (possibly all of it) fits next to `extra_bytes` of
// caller staging; picks the chunk size stored in the flow.
int general_plan_workspace(Flow& f, int op, int64_t N, size_t extra_bytes);
int general_run(Flow& f, const GeneralJob& job);
int general_backward_from_stash(Flow& f, const void* gy_host, const void* gld_host);
void general_release(Flow& f);
int base_sample_dev(Flow& f, int64_t N, uint64_t seed, void* z_dev);
int rqs_bin_search_host(int dtype, const void* knots_host, const void* v_host, int64_t M, int K, int32_t* bins_out);

}  // namespace nf
