// dataflow K1 instantiations for ode_model_id 3 (all band half-widths); see flow_kernel.cuh
#include "flow_kernel.cuh"
namespace magi {
cudaError_t launch_flow_model_3(const FlowArgs& a, int HB, int grid, size_t smem_bytes, cudaStream_t st) { return flow_launch_model<3>(a, HB, grid, smem_bytes, st); }
}
