// K1 instantiations for ode_model_id 1 (all band half-widths); see banded_kernel.cuh
#include "banded_kernel.cuh"
namespace magi {
cudaError_t launch_banded_model_1(const BandedArgs& a, int HB, int DW, size_t smem_bytes, cudaStream_t st) { return launch_model<1>(a, HB, DW, smem_bytes, st); }
}
