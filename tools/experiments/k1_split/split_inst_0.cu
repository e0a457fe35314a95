// K1-split instantiations for ode_model_id 0 (all band half-widths); see split_kernel.cuh
#include "split_kernel.cuh"
namespace magi {
cudaError_t launch_split_model_0(const SplitArgs& a, int HB, int sm_count, cudaStream_t st, long long* launches) { return split_launch_model<0>(a, HB, sm_count, st, launches); }
}
