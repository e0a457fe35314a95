// Batched FP64 DMMA GEMM (see gemm_f64.cuh).
#include "gemm_f64.cuh"

namespace magi {

constexpr int GEMM_BM = 64, GEMM_BN = 64, GEMM_BK = 16, GEMM_LD = 72;

__global__ void __launch_bounds__(256) gemm_f64_dmma_kernel(const GemmArgs g) {
    __shared__ double As[2][GEMM_BK][GEMM_LD];
    __shared__ double Bs[2][GEMM_BK][GEMM_LD];
    const int m0 = blockIdx.y * GEMM_BM, n0 = blockIdx.x * GEMM_BN;
    if (g.lower_only && n0 > m0 + GEMM_BM - 1) return;
    const int z1 = blockIdx.z % g.nb1, z2 = blockIdx.z / g.nb1;
    const double* A = g.A + z1 * g.bsA1 + z2 * g.bsA2;
    const double* B = g.B + z1 * g.bsB1 + z2 * g.bsB2;
    double* C = g.C + z1 * g.bsC1 + z2 * g.bsC2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, q = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    const bool a_m_contig = (g.rsA == 1), b_n_contig = (g.csB == 1);
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    int k_begin = 0;
    if (g.k_lo_from_tile) { int t = m0 > n0 ? m0 : n0; k_begin = (t / GEMM_BK) * GEMM_BK; }
    const int nkt = (g.K - k_begin + GEMM_BK - 1) / GEMM_BK;
    double ra[4], rb[4];
    auto gload = [&](int kt) {
        const int k0 = k_begin + kt * GEMM_BK;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = tid + 256 * j;
            int m, k;
            if (a_m_contig) { m = idx & 63; k = idx >> 6; } else { k = idx & 15; m = idx >> 4; }
            const int gm = m0 + m, gk = k0 + k;
            ra[j] = (gm < g.M && gk < g.K) ? A[gm * g.rsA + gk * g.csA] : 0.0;
            int n, kb;
            if (b_n_contig) { n = idx & 63; kb = idx >> 6; } else { kb = idx & 15; n = idx >> 4; }
            const int gn = n0 + n, gkb = k0 + kb;
            rb[j] = (gn < g.N && gkb < g.K) ? B[gkb * g.rsB + gn * g.csB] : 0.0;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = tid + 256 * j;
            int m, k;
            if (a_m_contig) { m = idx & 63; k = idx >> 6; } else { k = idx & 15; m = idx >> 4; }
            As[buf][k][m] = ra[j];
            int n, kb;
            if (b_n_contig) { n = idx & 63; kb = idx >> 6; } else { kb = idx & 15; n = idx >> 4; }
            Bs[buf][kb][n] = rb[j];
        }
    };
    if (nkt > 0) { gload(0); sstore(0); }
    __syncthreads();
    for (int kt = 0; kt < nkt; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nkt) gload(kt + 1);
#pragma unroll
        for (int k4 = 0; k4 < GEMM_BK / 4; ++k4) {
            double a[2], b[4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) a[mi] = As[cur][k4 * 4 + q][wm * 16 + mi * 8 + gid];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[cur][k4 * 4 + q][wn * 32 + ni * 8 + gid];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
        if (kt + 1 < nkt) sstore(cur ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = m0 + wm * 16 + mi * 8 + gid, c = n0 + wn * 32 + ni * 8 + 2 * q + u;
                if (r < g.M && c < g.N) {
                    double* p = C + r * g.rsC + c * g.csC;
                    const double v = g.alpha * acc[mi][ni][u];
                    *p = (g.beta == 0.0) ? v : v + g.beta * (*p);
                }
            }
}

cudaError_t launch_gemm(const GemmArgs& g, int batch, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0 || batch <= 0) return cudaSuccess;
    dim3 grid((g.N + GEMM_BN - 1) / GEMM_BN, (g.M + GEMM_BM - 1) / GEMM_BM, batch);
    gemm_f64_dmma_kernel<<<grid, 256, 0, st>>>(g);
    return cudaGetLastError();
}

}  // namespace magi
