// Temporary stubs (replaced as the setup / dense / sampler kernels land).
#include "magi_internal.cuh"
namespace magi {
void hmc_free(magi_handle*) {}
}
