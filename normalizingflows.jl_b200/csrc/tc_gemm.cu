// placeholder: filled in by the tcgen05 implementation
#include "tc_gemm.hpp"
namespace nf {
size_t tc_act_bytes(int64_t n, int width) { return (size_t)round_up(n, 128) * round_up(width, 64) * 4; }
size_t tc_weight_bytes(const Flow&) { return 0; }
int tc_prepare_weights(Flow&, const float*) { set_error("tcgen05 path not built"); return NF_ERR_UNSUPPORTED; }
int tc_gather_split(Flow&, const float*, int, const int*, int, int64_t, void*) { return NF_ERR_UNSUPPORTED; }
int tc_mlp_forward(Flow&, const LayerDesc&, int, int64_t, void*, std::vector<void*>&) { return NF_ERR_UNSUPPORTED; }
int tc_mlp_backward(Flow&, const LayerDesc&, int, int64_t, void*, std::vector<void*>&, float*, void*, float*, double*) { return NF_ERR_UNSUPPORTED; }
void tc_release(Flow&) {}
}
