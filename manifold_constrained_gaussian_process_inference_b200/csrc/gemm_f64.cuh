// Batched FP64 GEMM on DMMA.8x8x4 with generic element strides (any transpose / sub-matrix view), used by the GP
// setup (blocked Cholesky trailing updates, triangular inverse, m = C'Cinv, K = C'' - m C'^T) and by the dense
// (band = n-1) evaluation path.  C = alpha * A * B + beta * C, element A(i,k) = A[i*rsA + k*csA + z1*bsA1 + z2*bsA2].
#pragma once
#include "magi_common.cuh"

namespace magi {

struct GemmArgs {
    const double* A; long long rsA, csA, bsA1, bsA2;
    const double* B; long long rsB, csB, bsB1, bsB2;
    double* C;       long long rsC, csC, bsC1, bsC2;
    int M, N, K;
    int nb1;              // inner batch count: blockIdx.z = z1 + nb1 * z2
    double alpha, beta;
    int lower_only;       // skip output tiles that lie strictly above the diagonal (symmetric updates)
    int k_lo_from_tile;   // 1: A and B are lower-triangular-structured such that k < max(tile row0, tile col0) contributes 0
    int upper_only = 0;   // skip output tiles that lie strictly below the diagonal (only the upper triangle of C is wanted)
    int a_lower = 0;      // op(A)(m, k) is zero for k > m (lower-triangular A): k-tiles beyond the tile's last row are skipped
    int b_lower = 0;      // op(B)(k, n) is zero for k < n (lower-triangular B): k-tiles before the tile's first column are skipped
    int a_band = 0;       // > 0: A(m, k) is zero for |m - k| > a_band (band-truncated operator): k-tiles outside the band are skipped
    // optional stream-K work space (see gemm_f64.cu): partial tiles [kStreamKSlots][128 x 128] and one flag per slot; the caller
    // owns both, zero-initialises the flags once and passes a fresh non-zero epoch per launch.  Null: data-parallel tiles only.
    double* sk_work = nullptr;
    unsigned* sk_flags = nullptr;
    unsigned sk_epoch = 0;
    int sm_count = 0;     // SMs of the launch device (0: queried per launch)
    long long* extra_launches = nullptr;   // incremented when the call launches a second kernel (the skinny remainder rows)
};
constexpr int kStreamKSlots = 160;                       // >= SM count
constexpr size_t kStreamKWorkDoubles = (size_t)kStreamKSlots * 128 * 128;

cudaError_t launch_gemm(const GemmArgs& g, int batch, cudaStream_t st);

}  // namespace magi
