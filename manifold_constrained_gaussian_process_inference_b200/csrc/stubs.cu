// Temporary stubs (replaced as the setup / dense / sampler kernels land).
#include "magi_internal.cuh"
namespace magi {
int eval_dense_dev(magi_handle*, int, const double*, long long, double*, double*, cudaStream_t) { return set_error(MAGI_ERR_UNSUPPORTED, "dense mode not built yet"); }
void hmc_free(magi_handle*) {}
}
