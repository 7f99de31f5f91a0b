// Inverse-sweep fused elementwise kernels, double instantiations (see elementwise_impl.cuh).
#include "elementwise_impl.cuh"
namespace nf {
template int ew_run_dir<double, true>(Flow&, const Target*, const void*, int64_t, const void*, uint64_t, bool, void*, void*, void*, double*, bool);
template int ew_segment_dir<double, true>(Flow&, int, int, const void*, int64_t, const void*, void*, void*, bool, void*, const void*, double*);
}  // namespace nf
