// Persistent single-launch training kernels, double instantiations (see elementwise_impl.cuh).
#include "elementwise_impl.cuh"
namespace nf {
template int ew_train<double>(Flow&, const Target*, int64_t, uint64_t, int, int, double, double, double, double, void*, void*);
}  // namespace nf
