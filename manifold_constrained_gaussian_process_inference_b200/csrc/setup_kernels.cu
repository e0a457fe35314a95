// K3-K6: the one-off GP setup on device, replacing calculate_gp_covariances! (src/gaussian_process.jl:219-363):
//   K3  covariance build C, C', C''            (:249 kernelmatrix; :78-123 Matern-5/2; :128-154 RBF)
//   K4  blocked Cholesky of C+eI and of K+eI   (:295, :317; right-looking, DMMA GEMM trailing updates, pivot repair flag)
//   K5  inverse from the factor (triangular inverse by recursive doubling + L^-T L^-1), m = C' Cinv, K = C'' - m C'^T + eI
//       (:296, :302-307, :318) -- or, setup_mode "stable", W = L^-1 C'^T, K = C'' - W^T W + eI, m = (L^-T W)^T
//   K6  band extraction into diagonal-major tables (:358-360, mat2band :70-74)
// All matrices are column-major n x n, batched over the D dimensions (batch stride n*n).
#include <chrono>
#include <cmath>
#include <cfloat>
#include "magi_internal.cuh"
#include "gemm_f64.cuh"

namespace magi {

#define SCK(call, what) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_error(e__, what); } while (0)

constexpr int NBMAX = 64;

// ---------------- K3 ----------------
__global__ void cov_build_kernel(int kernel_id, int complexity, const double* __restrict__ tvec, const double* __restrict__ phi,
                                 int n, double* __restrict__ C, double* __restrict__ Cp, double* __restrict__ Cpp) {
    const int d = blockIdx.z;
    const double var = phi[2 * d], l = phi[2 * d + 1];
    const size_t nn = (size_t)n * n;
    const double sqrt5 = sqrt(5.0);
    const double s = 1.0 / l;                       // ScaleTransform(1/l): inputs are scaled first (kernels.jl:49,80)
    const double l_sq = l * l, l_cub = l * l * l, l_quad = l_sq * l_sq;
    const double t3sq = 1.0 / (3.0 * l_sq), t3cub = 1.0 / (3.0 * l_cub);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < nn; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n), j = (int)(idx / n);     // column-major: element (i, j)
        const double ti = tvec[i], tj = tvec[j];
        double c, cp = 0.0, cpp = 0.0;
        if (kernel_id == MAGI_KERNEL_MATERN52) {
            const double r = fabs(ti * s - tj * s);
            c = var * ((1.0 + sqrt5 * r + 5.0 * r * r / 3.0) * exp(-sqrt5 * r));
            if (complexity >= 2) {
                if (i == j) { cp = 0.0; cpp = 5.0 * var / (3.0 * l_sq); }                      // :103-105
                else {
                    const double t_diff = ti - tj, dist = fabs(t_diff), dist_sq = dist * dist;
                    const double sgn = (t_diff > 0) - (t_diff < 0);
                    const double ex = exp(-sqrt5 * dist / l);
                    const double common = ex * (5 * dist * t3sq + 5 * sqrt5 * dist_sq * t3cub);  // :113
                    cp = -sgn * var * common;                                                     // :114
                    const double term1 = (-sqrt5 / l * ex) * (5 * dist * t3sq + 5 * sqrt5 * dist_sq * t3cub);   // :116
                    const double term2 = ex * (5 * t3sq + 10 * sqrt5 * dist * t3cub);                         // :117
                    cpp = var * (term1 + term2);                                                  // :118
                }
            }
        } else if (kernel_id == MAGI_KERNEL_MATERN_NU12 || kernel_id == MAGI_KERNEL_MATERN_NU32 || kernel_id == MAGI_KERNEL_MATERN_NU52) {
            // MaternKernel(nu) (kernels.jl:109-118) in closed form; no analytic derivatives in the reference (:278-280)
            const double r = fabs(ti * s - tj * s);
            if (kernel_id == MAGI_KERNEL_MATERN_NU12) c = var * exp(-r);
            else if (kernel_id == MAGI_KERNEL_MATERN_NU32) c = var * ((1.0 + sqrt(3.0) * r) * exp(-sqrt(3.0) * r));
            else c = var * ((1.0 + sqrt5 * r + 5.0 * r * r / 3.0) * exp(-sqrt5 * r));
        } else {
            const double dd = ti * s - tj * s;
            c = var * exp(-(dd * dd) / 2.0);
            if (complexity >= 2) {
                const double t_diff = ti - tj;
                cp = -c * t_diff / l_sq;                                   // :149
                cpp = c * (1.0 / l_sq - (t_diff * t_diff) / l_quad);       // :150
            }
        }
        C[d * nn + idx] = c;
        Cp[d * nn + idx] = cp;
        Cpp[d * nn + idx] = cpp;
    }
}

// dst = src + eps*I ; optionally symmetrise from the UPPER triangle (Symmetric(.) reads the upper triangle, :257, :306)
__global__ void add_jitter_sym_kernel(const double* __restrict__ src, double* __restrict__ dst, int n, double eps, int sym_upper) {
    const int d = blockIdx.z;
    const size_t nn = (size_t)n * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < nn; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n), j = (int)(idx / n);
        double v;
        if (sym_upper && i > j) v = src[d * nn + (size_t)i * n + j];   // element (j, i) of the upper triangle
        else v = src[d * nn + idx];
        if (i == j) v += eps;
        dst[d * nn + idx] = v;
    }
}

__global__ void mirror_lower_kernel(double* __restrict__ A, int n) {       // A(j,i) = A(i,j) for i > j
    const int d = blockIdx.z;
    const size_t nn = (size_t)n * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < nn; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n), j = (int)(idx / n);
        if (i < j) A[d * nn + idx] = A[d * nn + (size_t)i * n + j];
    }
}

__global__ void fill_kernel(double* __restrict__ A, size_t count, double v) {
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < count; idx += (size_t)gridDim.x * blockDim.x) A[idx] = v;
}

__global__ void set_diag_kernel(double* __restrict__ A, int n, double v) {   // A = v*I (batched), off-diagonal zeroed
    const int d = blockIdx.z;
    const size_t nn = (size_t)n * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < nn; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n), j = (int)(idx / n);
        A[d * nn + idx] = (i == j) ? v : 0.0;
    }
}

// per batch item: stats[2d] = max |diag|, stats[2d+1] = 1 if any off-diagonal... (used for: pivot tolerance; "all zero" test)
// (grid = (D, chunks): every block reduces a slice and merges with an atomic max on the bit pattern -- non-negative doubles
// order like unsigned integers; `out` must be zeroed before the launch)
__global__ void absmax_kernel(const double* __restrict__ A, int n, int diag_only, double* __restrict__ out) {
    const int d = blockIdx.x;
    const size_t nn = (size_t)n * n;
    __shared__ double red[256];
    double m = 0.0;
    const size_t tid = (size_t)blockIdx.y * blockDim.x + threadIdx.x, stride = (size_t)gridDim.y * blockDim.x;
    if (diag_only) { for (size_t i = tid; i < (size_t)n; i += stride) m = fmax(m, fabs(A[d * nn + i * n + i])); }
    else { for (size_t i = tid; i < nn; i += stride) m = fmax(m, fabs(A[d * nn + i])); }
    red[threadIdx.x] = m;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
    if (threadIdx.x == 0) {
        const double v = red[0];
        // (fmax drops NaNs, as in a sequential max of |.|: v is a non-negative number)
        atomicMax(reinterpret_cast<unsigned long long*>(out + d), (unsigned long long)__double_as_longlong(v));
    }
}

// ---------------- K4: diagonal block factorisation + its triangular inverse ----------------
// One block per batch item.  Factors the nb x nb diagonal block at (j0, j0) of A (lower triangle valid) in shared memory,
// writes L_jj back (strict upper part of the block zeroed) and inv(L_jj) to Tinv[d][blk] (NBMAX x NBMAX, column-major, lower).
// A non-positive pivot does not abort: it is replaced like the reference's cholesky(Positive, .) is recalled to do
// (|pivot|, or the tolerance when tiny) and counted in repaired[d] (unverified rule).
__global__ void __launch_bounds__(256) potf2_inv_kernel(double* __restrict__ A, int n, int j0, int nb, double* __restrict__ Tinv,
                                                        int blk, int nblk, const double* __restrict__ diagmax, int* __restrict__ repaired) {
    __shared__ double S[NBMAX][NBMAX + 1];   // lower: L; strict upper (transposed): inv(L) below its diagonal
    __shared__ double xd[NBMAX];             // diagonal of inv(L)
    const int d = blockIdx.x;
    double* Ad = A + (size_t)d * n * n;
    const double tol = (double)n * DBL_EPSILON * diagmax[d];
    for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
        const int i = idx % nb, j = idx / nb;
        S[i][j] = (i >= j) ? Ad[(size_t)(j0 + j) * n + (j0 + i)] : 0.0;
    }
    __syncthreads();
    // Right-looking factorisation with ONE block barrier per column: every thread derives the (repaired) pivot p_j itself from
    // S[j][j] and applies the rank-1 update with the UNSCALED column, S[i][k] -= S[i][j] S[k][j] / p_j; the columns are scaled by
    // 1 / sqrt(p_j) in one parallel pass afterwards.  (Round 1: thread 0 computed the pivot, three barriers per column.)
    __shared__ double piv[NBMAX];
    for (int j = 0; j < nb; ++j) {
        double p = S[j][j];
        if (!(p > tol)) {
            if (threadIdx.x == 0) atomicAdd(&repaired[d], 1);
            p = (fabs(p) > tol) ? fabs(p) : (tol > 0 ? tol : 1.0);
        }
        if (threadIdx.x == 0) piv[j] = p;
        const double inv_p = 1.0 / p;
        const int m = nb - j - 1;
        for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
            const int ii = j + 1 + idx % m, jj = j + 1 + idx / m;
            if (ii >= jj) S[ii][jj] -= S[ii][j] * S[jj][j] * inv_p;
        }
        __syncthreads();
    }
    for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
        const int i = idx % nb, j = idx / nb;
        if (i > j) S[i][j] /= sqrt(piv[j]);
    }
    __syncthreads();
    if (threadIdx.x < nb) S[threadIdx.x][threadIdx.x] = sqrt(piv[threadIdx.x]);
    __syncthreads();
    // inverse of the lower-triangular block: thread c solves column c by forward substitution; X(i,c), i > c, is kept
    // at S[c][i] (the unused strict upper triangle), X(c,c) in xd[c]
    if (threadIdx.x < nb) {
        const int c = threadIdx.x;
        const double xc = 1.0 / S[c][c];
        xd[c] = xc;
        for (int i = c + 1; i < nb; ++i) {
            double s = S[i][c] * xc;
            for (int k = c + 1; k < i; ++k) s += S[i][k] * S[c][k];
            S[c][i] = -s / S[i][i];
        }
    }
    __syncthreads();
    double* T = Tinv + ((size_t)d * nblk + blk) * NBMAX * NBMAX;
    for (int idx = threadIdx.x; idx < NBMAX * NBMAX; idx += blockDim.x) {
        const int i = idx % NBMAX, j = idx / NBMAX;
        double x = 0.0;
        if (i < nb && j < nb) x = (i == j) ? xd[i] : (i > j ? S[j][i] : 0.0);
        T[idx] = x;
        if (i < nb && j < nb) Ad[(size_t)(j0 + j) * n + (j0 + i)] = (i >= j) ? S[i][j] : 0.0;
    }
}

// zero the strict upper triangle (after the factorisation the upper part still holds the input)
__global__ void zero_upper_kernel(double* __restrict__ A, int n) {
    const int d = blockIdx.z;
    const size_t nn = (size_t)n * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < nn; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n), j = (int)(idx / n);
        if (i < j) A[d * nn + idx] = 0.0;
    }
}

// scatter the inverted diagonal blocks into X (n x n, zero elsewhere)
__global__ void scatter_tinv_kernel(const double* __restrict__ Tinv, double* __restrict__ X, int n, int NB, int nblk) {
    const int d = blockIdx.z;
    const size_t nn = (size_t)n * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < nn; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n), j = (int)(idx / n);
        const int bi = i / NB, bj = j / NB;
        double v = 0.0;
        if (bi == bj) v = Tinv[((size_t)d * nblk + bi) * NBMAX * NBMAX + (size_t)(j - bj * NB) * NBMAX + (i - bi * NB)];
        X[d * nn + idx] = v;
    }
}

// ---------------- K6 ----------------
__global__ void band_extract_kernel(const double* __restrict__ dense, double* __restrict__ band, int n, int b) {
    const int d = blockIdx.z;
    const size_t tab = (size_t)(2 * b + 1) * n, nn = (size_t)n * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < tab; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n), off = (int)(idx / n) - b;
        const int j = i + off;
        band[d * tab + idx] = (j >= 0 && j < n) ? dense[d * nn + (size_t)j * n + i] : 0.0;    // A[i][j], column-major
    }
}

// ------------------------------------------------------------------------------------------------------------
struct SetupCtx {
    int n, D, b, kernel_id, mode, complexity;
    double jitter;
    const double* d_tvec;
    const double* d_phi;
    double* dense[7];      // C, Cinv, Cprime, Cdoubleprime, mphi, Kphi, Kinv (each D*n*n)
    double* band[3];       // CinvBand, mphiBand, KinvBand (each D*(2b+1)*n)
    cudaStream_t st;
    long long launches = 0;
    std::vector<int> rep_c, rep_k;
    double alloc_ms = 0.0, kernel_ms = 0.0;   // host time spent in cudaMalloc / device time from the covariance build to the band tables
};

static dim3 ew_grid(size_t count, int D) {
    size_t blocks = (count + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    return dim3((unsigned)blocks, 1, (unsigned)D);
}

static GemmArgs gemm_cm(const double* A, bool tA, const double* B, bool tB, double* C, int M, int N, int K, int n, double alpha, double beta) {
    // column-major n x n matrices batched by n*n; op(A)(i,k), op(B)(k,j)
    GemmArgs g{};
    g.A = A; g.rsA = tA ? n : 1; g.csA = tA ? 1 : n; g.bsA1 = (long long)n * n; g.bsA2 = 0;
    g.B = B; g.rsB = tB ? n : 1; g.csB = tB ? 1 : n; g.bsB1 = (long long)n * n; g.bsB2 = 0;
    g.C = C; g.rsC = 1; g.csC = n; g.bsC1 = (long long)n * n; g.bsC2 = 0;
    g.M = M; g.N = N; g.K = K; g.nb1 = 1 << 30; g.alpha = alpha; g.beta = beta; g.lower_only = 0; g.k_lo_from_tile = 0;
    return g;
}

// Blocked right-looking Cholesky of the (lower-valid) matrices in A (in place; strict upper zeroed afterwards), and
// X = inv(L) by recursive doubling.  T is an n x n work buffer per batch item.
static int chol_and_inverse(SetupCtx& c, double* A, double* X, double* T, double* d_tinv, double* d_stat, int* d_rep, int NB, int nblk) {
    const int n = c.n, D = c.D;
    SCK(cudaMemsetAsync(d_stat, 0, sizeof(double) * D, c.st), "memset stats");
    absmax_kernel<<<dim3(D, 1), 256, 0, c.st>>>(A, n, 1, d_stat); c.launches++;
    SCK(cudaMemsetAsync(d_rep, 0, sizeof(int) * D, c.st), "memset repaired");
    const int NBO = (n >= 8 * NB) ? 4 * NB : NB;         // outer panel width
    for (int blk = 0; blk < nblk; ++blk) {
        const int j0 = blk * NB, nb = (n - j0 < NB) ? n - j0 : NB;
        potf2_inv_kernel<<<D, 256, 0, c.st>>>(A, n, j0, nb, d_tinv, blk, nblk, d_stat, d_rep); c.launches++;
        SCK(cudaGetLastError(), "potf2_inv_kernel");
        const int rem = n - j0 - nb;
        if (rem > 0) {
            // panel: A[j0+nb:, j0:j0+nb] <- A[j0+nb:, j0:j0+nb] * inv(L_jj)^T   (one 64-wide tile column: in-place safe)
            GemmArgs g{};
            g.A = A + (size_t)j0 * n + (j0 + nb); g.rsA = 1; g.csA = n; g.bsA1 = (long long)n * n; g.bsA2 = 0;
            g.B = d_tinv + (size_t)blk * NBMAX * NBMAX; g.rsB = NBMAX; g.csB = 1; g.bsB1 = (long long)nblk * NBMAX * NBMAX; g.bsB2 = 0;   // B(k,j) = Tinv(j,k)
            g.C = A + (size_t)j0 * n + (j0 + nb); g.rsC = 1; g.csC = n; g.bsC1 = (long long)n * n; g.bsC2 = 0;
            g.M = rem; g.N = nb; g.K = nb; g.nb1 = 1 << 30; g.alpha = 1.0; g.beta = 0.0;
            SCK(launch_gemm(g, D, c.st), "panel gemm"); c.launches++;
            // trailing update, two levels: inside the current outer panel (NBO columns) the update follows every block
            // (K = nb); the rest of the matrix is updated once per outer panel with K = NBO -- a quarter of the
            // read-modify-write passes over the trailing matrix, and GEMMs deep enough to run on the tensor pipe
            // instead of on HBM.  Lower tiles only.
            const int j1 = j0 + nb, jend = (j0 / NBO + 1) * NBO < n ? (j0 / NBO + 1) * NBO : n, jbeg = (j0 / NBO) * NBO;
            if (jend > j1) {
                GemmArgs u{};
                u.A = A + (size_t)j0 * n + j1; u.rsA = 1; u.csA = n; u.bsA1 = (long long)n * n;
                u.B = A + (size_t)j0 * n + j1; u.rsB = n; u.csB = 1; u.bsB1 = (long long)n * n;      // B(k,j) = P(j,k)
                u.C = A + (size_t)j1 * n + j1; u.rsC = 1; u.csC = n; u.bsC1 = (long long)n * n;
                u.M = rem; u.N = jend - j1; u.K = nb; u.nb1 = 1 << 30; u.alpha = -1.0; u.beta = 1.0; u.lower_only = 1;
                SCK(launch_gemm(u, D, c.st), "trailing gemm (panel)"); c.launches++;
            }
            if (j1 == jend && jend < n) {
                GemmArgs u{};
                u.A = A + (size_t)jbeg * n + jend; u.rsA = 1; u.csA = n; u.bsA1 = (long long)n * n;
                u.B = A + (size_t)jbeg * n + jend; u.rsB = n; u.csB = 1; u.bsB1 = (long long)n * n;
                u.C = A + (size_t)jend * n + jend; u.rsC = 1; u.csC = n; u.bsC1 = (long long)n * n;
                u.M = n - jend; u.N = n - jend; u.K = jend - jbeg; u.nb1 = 1 << 30; u.alpha = -1.0; u.beta = 1.0; u.lower_only = 1;
                SCK(launch_gemm(u, D, c.st), "trailing gemm (outer)"); c.launches++;
            }
        }
    }
    zero_upper_kernel<<<ew_grid((size_t)n * n, D), 256, 0, c.st>>>(A, n); c.launches++;
    // X = blockdiag(inv(L_jj)); then combine pairs of size s: X21 = -X22 * L21 * X11
    scatter_tinv_kernel<<<ew_grid((size_t)n * n, D), 256, 0, c.st>>>(d_tinv, X, n, NB, nblk); c.launches++;
    for (long long s = NB; s < n; s *= 2) {
        const int npairs = (int)((n - s + 2 * s - 1) / (2 * s));
        for (int pass = 0; pass < 2; ++pass) {
            // pass 0: all pairs with a full second block; pass 1: the ragged last pair (if any)
            const long long last_m2 = n - ((long long)(npairs - 1) * 2 * s + s);
            const bool ragged = last_m2 < s;
            int p0, cnt, m2;
            if (pass == 0) { p0 = 0; cnt = ragged ? npairs - 1 : npairs; m2 = (int)s; }
            else { if (!ragged) break; p0 = npairs - 1; cnt = 1; m2 = (int)last_m2; }
            if (cnt <= 0 || m2 <= 0) continue;
            const long long pair_stride = 2 * s * n + 2 * s;           // next pair along the diagonal (column-major)
            const long long o11 = (long long)p0 * pair_stride;          // (2ps, 2ps)
            const long long o21 = o11 + s;                              // (2ps+s, 2ps)
            const long long o22 = o11 + s * n + s;                      // (2ps+s, 2ps+s)
            GemmArgs g1{};   // T21 = L21 * X11     (m2 x s) = (m2 x s)(s x s)
            g1.A = A + o21; g1.rsA = 1; g1.csA = n; g1.bsA1 = pair_stride; g1.bsA2 = (long long)n * n;
            g1.B = X + o11; g1.rsB = 1; g1.csB = n; g1.bsB1 = pair_stride; g1.bsB2 = (long long)n * n;
            g1.C = T + o21; g1.rsC = 1; g1.csC = n; g1.bsC1 = pair_stride; g1.bsC2 = (long long)n * n;
            g1.M = m2; g1.N = (int)s; g1.K = (int)s; g1.nb1 = cnt; g1.alpha = 1.0; g1.beta = 0.0; g1.b_lower = 1;   // X11 is lower triangular
            SCK(launch_gemm(g1, cnt * D, c.st), "trinv gemm 1"); c.launches++;
            GemmArgs g2{};   // X21 = -X22 * T21    (m2 x s) = (m2 x m2)(m2 x s)
            g2.A = X + o22; g2.rsA = 1; g2.csA = n; g2.bsA1 = pair_stride; g2.bsA2 = (long long)n * n;
            g2.B = T + o21; g2.rsB = 1; g2.csB = n; g2.bsB1 = pair_stride; g2.bsB2 = (long long)n * n;
            g2.C = X + o21; g2.rsC = 1; g2.csC = n; g2.bsC1 = pair_stride; g2.bsC2 = (long long)n * n;
            g2.M = m2; g2.N = (int)s; g2.K = m2; g2.nb1 = cnt; g2.alpha = -1.0; g2.beta = 0.0; g2.a_lower = 1;   // so is X22
            SCK(launch_gemm(g2, cnt * D, c.st), "trinv gemm 2"); c.launches++;
        }
    }
    return MAGI_OK;
}

// inv = X^T X (X = inv(L), lower triangular), lower tiles then mirrored: exactly symmetric like potri + copytri (:296, :318)
static int inverse_from_factor(SetupCtx& c, const double* X, double* inv) {
    const int n = c.n;
    GemmArgs g = gemm_cm(X, true, X, false, inv, n, n, n, n, 1.0, 0.0);
    g.lower_only = 1; g.k_lo_from_tile = 1;
    SCK(launch_gemm(g, c.D, c.st), "potri gemm"); c.launches++;
    mirror_lower_kernel<<<ew_grid((size_t)n * n, c.D), 256, 0, c.st>>>(inv, n); c.launches++;
    return MAGI_OK;
}

int gp_setup_device(SetupCtx& c) {
    const int n = c.n, D = c.D;
    const size_t nn = (size_t)n * n;
    double *C = c.dense[MAGI_MAT_C], *Cinv = c.dense[MAGI_MAT_CINV], *Cp = c.dense[MAGI_MAT_CPRIME], *Cpp = c.dense[MAGI_MAT_CDOUBLEPRIME];
    double *mphi = c.dense[MAGI_MAT_MPHI], *Kphi = c.dense[MAGI_MAT_KPHI], *Kinv = c.dense[MAGI_MAT_KINV];
    int levels = 0;
    while (((long long)NBMAX << levels) < n) ++levels;
    int NB = (n + (1 << levels) - 1) >> levels;
    if (NB < 1) NB = 1;
    const int nblk = (n + NB - 1) / NB;
    double *L = nullptr, *X = nullptr, *T = nullptr, *d_tinv = nullptr, *d_stat = nullptr;
    int* d_rep = nullptr;
    auto cleanup = [&]() { cudaFree(L); cudaFree(X); cudaFree(T); cudaFree(d_tinv); cudaFree(d_stat); cudaFree(d_rep); };
#define SETUP_TRY(expr) do { int rc__ = (expr); if (rc__ != MAGI_OK) { cleanup(); return rc__; } } while (0)
#define SETUP_CUDA(call, what) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return cuda_error(e__, what); } } while (0)
    const auto t_alloc0 = std::chrono::steady_clock::now();
    SETUP_CUDA(cudaMalloc(&L, sizeof(double) * nn * D), "cudaMalloc setup L");
    SETUP_CUDA(cudaMalloc(&X, sizeof(double) * nn * D), "cudaMalloc setup X");
    SETUP_CUDA(cudaMalloc(&T, sizeof(double) * nn * D), "cudaMalloc setup T");
    SETUP_CUDA(cudaMalloc(&d_tinv, sizeof(double) * (size_t)D * nblk * NBMAX * NBMAX), "cudaMalloc tinv");
    SETUP_CUDA(cudaMalloc(&d_stat, sizeof(double) * 2 * D), "cudaMalloc stat");
    SETUP_CUDA(cudaMalloc(&d_rep, sizeof(int) * D), "cudaMalloc rep");
    c.alloc_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_alloc0).count();
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);
    cudaEventRecord(ev0, c.st);
    SETUP_CUDA(cudaMemsetAsync(T, 0, sizeof(double) * nn * D, c.st), "memset T");
    const dim3 eg = ew_grid(nn, D);
    const bool want_deriv = c.complexity >= 2 && (c.kernel_id == MAGI_KERNEL_MATERN52 || c.kernel_id == MAGI_KERNEL_RBF);
    cov_build_kernel<<<eg, 256, 0, c.st>>>(c.kernel_id, want_deriv ? 2 : 0, c.d_tvec, c.d_phi, n, C, Cp, Cpp); c.launches++;
    SETUP_CUDA(cudaGetLastError(), "cov_build_kernel");
    // derivatives "all zero" test of the reference (:299), decided PER DIMENSION like the reference's per-GPCov call: a
    // dimension with all-zero C' or C'' (n = 1; a lengthscale so small that every off-diagonal entry underflows) takes the
    // fallback, the others the full path
    std::vector<double> st(2 * D, 0.0);
    SETUP_CUDA(cudaMemsetAsync(d_stat, 0, sizeof(double) * 2 * D, c.st), "memset stats");
    const int abs_chunks = (int)std::min<size_t>(64, ((size_t)n * n + 65535) / 65536);
    absmax_kernel<<<dim3(D, abs_chunks), 256, 0, c.st>>>(Cp, n, 0, d_stat); c.launches++;
    absmax_kernel<<<dim3(D, abs_chunks), 256, 0, c.st>>>(Cpp, n, 0, d_stat + D); c.launches++;
    SETUP_CUDA(cudaMemcpyAsync(st.data(), d_stat, sizeof(double) * 2 * D, cudaMemcpyDeviceToHost, c.st), "D2H stats");
    SETUP_CUDA(cudaStreamSynchronize(c.st), "sync");
    std::vector<char> deriv_d(D, 0);
    bool deriv = false;
    for (int d = 0; d < D; ++d) { deriv_d[d] = want_deriv && st[d] > 0.0 && st[D + d] > 0.0; deriv = deriv || deriv_d[d]; }
    // C + eI -> L ; X = inv(L) ; Cinv
    add_jitter_sym_kernel<<<eg, 256, 0, c.st>>>(C, L, n, c.jitter, 1); c.launches++;
    SETUP_TRY(chol_and_inverse(c, L, X, T, d_tinv, d_stat, d_rep, NB, nblk));
    c.rep_c.assign(D, 0);
    SETUP_CUDA(cudaMemcpyAsync(c.rep_c.data(), d_rep, sizeof(int) * D, cudaMemcpyDeviceToHost, c.st), "D2H repaired");
    SETUP_TRY(inverse_from_factor(c, X, Cinv));
    if (deriv) {
        if (c.mode == MAGI_SETUP_REFERENCE_ORDER) {
            GemmArgs g = gemm_cm(Cp, false, Cinv, false, mphi, n, n, n, n, 1.0, 0.0);            // m = C' Cinv  (:302)
            SETUP_CUDA(launch_gemm(g, D, c.st), "mphi gemm"); c.launches++;
            SETUP_CUDA(cudaMemcpyAsync(T, Cpp, sizeof(double) * nn * D, cudaMemcpyDeviceToDevice, c.st), "copy Cpp");
            GemmArgs k = gemm_cm(mphi, false, Cp, true, T, n, n, n, n, -1.0, 1.0);               // C'' - m C'^T (:304)
            k.upper_only = 1;                                  // Symmetric(.) reads the upper triangle only (:306)
            SETUP_CUDA(launch_gemm(k, D, c.st), "Kphi gemm"); c.launches++;
        } else {
            // stable: W = inv(L) C'^T  (X is the explicit triangular inverse; the products are Gram forms)
            GemmArgs w = gemm_cm(X, false, Cp, true, Kinv /*scratch: W*/, n, n, n, n, 1.0, 0.0);
            w.a_lower = 1;                                     // X = inv(L) is lower triangular: half of the k-tiles are zero
            SETUP_CUDA(launch_gemm(w, D, c.st), "W gemm"); c.launches++;
            SETUP_CUDA(cudaMemcpyAsync(T, Cpp, sizeof(double) * nn * D, cudaMemcpyDeviceToDevice, c.st), "copy Cpp");
            GemmArgs k = gemm_cm(Kinv, true, Kinv, false, T, n, n, n, n, -1.0, 1.0);             // C'' - W^T W
            k.upper_only = 1;                                  // Symmetric(.) reads the upper triangle only
            SETUP_CUDA(launch_gemm(k, D, c.st), "Kphi syrk"); c.launches++;
            GemmArgs m = gemm_cm(Kinv, true, X, false, mphi, n, n, n, n, 1.0, 0.0);              // m = W^T inv(L) = C' inv(L)^T inv(L)
            m.b_lower = 1;                                     // inv(L)(k, j) = 0 for k < j
            SETUP_CUDA(launch_gemm(m, D, c.st), "mphi gemm"); c.launches++;
        }
        add_jitter_sym_kernel<<<eg, 256, 0, c.st>>>(T, Kphi, n, c.jitter, 1); c.launches++;       // Symmetric(K + eI) from the upper triangle (:306-307)
        SETUP_CUDA(cudaMemcpyAsync(L, Kphi, sizeof(double) * nn * D, cudaMemcpyDeviceToDevice, c.st), "copy Kphi");
        SETUP_TRY(chol_and_inverse(c, L, X, T, d_tinv, d_stat, d_rep, NB, nblk));
        c.rep_k.assign(D, 0);
        SETUP_CUDA(cudaMemcpyAsync(c.rep_k.data(), d_rep, sizeof(int) * D, cudaMemcpyDeviceToHost, c.st), "D2H repaired");
        SETUP_TRY(inverse_from_factor(c, X, Kinv));
        SETUP_CUDA(cudaStreamSynchronize(c.st), "sync");
        for (int d = 0; d < D; ++d) {
            if (deriv_d[d]) continue;                 // this dimension alone takes the fallback (:319-331)
            const dim3 e1 = ew_grid(nn, 1);
            SETUP_CUDA(cudaMemsetAsync(Cp + d * nn, 0, sizeof(double) * nn, c.st), "memset");
            SETUP_CUDA(cudaMemsetAsync(Cpp + d * nn, 0, sizeof(double) * nn, c.st), "memset");
            SETUP_CUDA(cudaMemsetAsync(mphi + d * nn, 0, sizeof(double) * nn, c.st), "memset");
            set_diag_kernel<<<e1, 256, 0, c.st>>>(Kphi + d * nn, n, c.jitter); c.launches++;
            set_diag_kernel<<<e1, 256, 0, c.st>>>(Kinv + d * nn, n, 1.0 / c.jitter); c.launches++;
            c.rep_k[d] = 0;
        }
    } else {
        // zero-derivative fallback (:278-280, :319-331): C' = C'' = m = 0, K = eI, Kinv = I/e
        SETUP_CUDA(cudaMemsetAsync(Cp, 0, sizeof(double) * nn * D, c.st), "memset");
        SETUP_CUDA(cudaMemsetAsync(Cpp, 0, sizeof(double) * nn * D, c.st), "memset");
        SETUP_CUDA(cudaMemsetAsync(mphi, 0, sizeof(double) * nn * D, c.st), "memset");
        set_diag_kernel<<<eg, 256, 0, c.st>>>(Kphi, n, c.jitter); c.launches++;
        set_diag_kernel<<<eg, 256, 0, c.st>>>(Kinv, n, 1.0 / c.jitter); c.launches++;
        c.rep_k.assign(D, 0);
    }
    const dim3 bg = ew_grid((size_t)(2 * c.b + 1) * n, D);
    band_extract_kernel<<<bg, 256, 0, c.st>>>(Cinv, c.band[0], n, c.b); c.launches++;
    band_extract_kernel<<<bg, 256, 0, c.st>>>(mphi, c.band[1], n, c.b); c.launches++;
    band_extract_kernel<<<bg, 256, 0, c.st>>>(Kinv, c.band[2], n, c.b); c.launches++;
    SETUP_CUDA(cudaGetLastError(), "band_extract_kernel");
    cudaEventRecord(ev1, c.st);
    SETUP_CUDA(cudaStreamSynchronize(c.st), "setup sync");
    { float ms = 0.f; if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) c.kernel_ms = ms; }
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    cleanup();
    return MAGI_OK;
}

// setup for a whole handle (all D dimensions batched)
int run_device_setup(magi_handle* h) {
    const size_t nn = (size_t)h->n * h->n;
    const auto t_alloc0 = std::chrono::steady_clock::now();
    for (int i = 0; i < 7; ++i) {
        cudaError_t e = cudaMalloc(&h->d_dense[i], sizeof(double) * nn * h->D);
        if (e != cudaSuccess) return cuda_error(e, "cudaMalloc dense GP matrices");
    }
    const double alloc_dense_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_alloc0).count();
    double *d_t = nullptr, *d_phi = nullptr;
    if (cudaMalloc(&d_t, sizeof(double) * h->n) != cudaSuccess || cudaMalloc(&d_phi, sizeof(double) * 2 * h->D) != cudaSuccess) {
        cudaFree(d_t); cudaFree(d_phi);
        return set_error(MAGI_ERR_CUDA, "cudaMalloc tvec/phi failed");
    }
    cudaMemcpyAsync(d_t, h->tvec.data(), sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_phi, h->phi.data(), sizeof(double) * 2 * h->D, cudaMemcpyHostToDevice, h->stream);
    SetupCtx c;
    c.n = h->n; c.D = h->D; c.b = h->b; c.kernel_id = h->kernel_id; c.mode = h->setup_mode; c.complexity = 2; c.jitter = h->jitter;
    c.d_tvec = d_t; c.d_phi = d_phi; c.st = h->stream;
    for (int i = 0; i < 7; ++i) c.dense[i] = h->d_dense[i];
    for (int i = 0; i < 3; ++i) c.band[i] = h->d_band[i];
    int rc = gp_setup_device(c);
    cudaFree(d_t); cudaFree(d_phi);
    h->launches += c.launches;
    if (rc != MAGI_OK) return rc;
    h->setup_alloc_ms = alloc_dense_ms + c.alloc_ms; h->setup_kernel_ms = c.kernel_ms;
    h->repaired_c = c.rep_c; h->repaired_k = c.rep_k;
    std::fill(h->band_set.begin(), h->band_set.end(), 1);
    h->tables_ready = true; h->frag_dirty = true; h->frag_nat_dirty = true; h->frag_bp_dirty = true; h->steptab_dirty = true; h->dense_band_dirty = true;
    return MAGI_OK;
}

}  // namespace magi

using namespace magi;

// Stand-alone GPCov computation for one dimension (the reference's calculate_gp_covariances!, gaussian_process.jl:219-363):
// host in (kernel, phi = [variance, lengthscale], tvec), host out (7 dense n x n column-major + 3 band tables).
extern "C" int magi_gp_covariances(int kernel_id, const double* phi, const double* tvec, int n, int bandsize, double jitter,
                                   int complexity, int setup_mode, int device, double* C, double* Cinv, double* Cprime,
                                   double* Cdoubleprime, double* mphi, double* Kphi, double* Kinv, double* CinvBand,
                                   double* mphiBand, double* KinvBand, int* repaired /* [2] or NULL */) {
    if (!phi || !tvec || n < 1) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_gp_covariances: bad argument");
    if (setup_mode != MAGI_SETUP_REFERENCE_ORDER && setup_mode != MAGI_SETUP_STABLE) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_gp_covariances: bad setup_mode");
    if (kernel_id < MAGI_KERNEL_MATERN52 || kernel_id > MAGI_KERNEL_MATERN_NU52) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_gp_covariances: unknown kernel_id");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_error(MAGI_ERR_CUDA, "magi_gp_covariances: no CUDA device available (no CPU fallback)");
    if (cudaSetDevice(device) != cudaSuccess) return set_error(MAGI_ERR_CUDA, "cudaSetDevice failed");
    int b = bandsize; if (b > n - 1) b = n - 1; if (b < 0) b = 0;
    const size_t nn = (size_t)n * n, tab = (size_t)(2 * b + 1) * n;
    SetupCtx c;
    c.n = n; c.D = 1; c.b = b; c.kernel_id = kernel_id; c.mode = setup_mode; c.complexity = complexity; c.jitter = jitter;
    double* pool = nullptr;
    if (cudaMalloc(&pool, sizeof(double) * (7 * nn + 3 * tab + n + 2)) != cudaSuccess) return set_error(MAGI_ERR_CUDA, "cudaMalloc failed");
    for (int i = 0; i < 7; ++i) c.dense[i] = pool + i * nn;
    for (int i = 0; i < 3; ++i) c.band[i] = pool + 7 * nn + i * tab;
    double* d_t = pool + 7 * nn + 3 * tab; double* d_phi = d_t + n;
    cudaStream_t st; cudaStreamCreate(&st); c.st = st;
    cudaMemcpyAsync(d_t, tvec, sizeof(double) * n, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_phi, phi, sizeof(double) * 2, cudaMemcpyHostToDevice, st);
    c.d_tvec = d_t; c.d_phi = d_phi;
    int rc = gp_setup_device(c);
    if (rc == MAGI_OK) {
        double* outs[7] = {C, Cinv, Cprime, Cdoubleprime, mphi, Kphi, Kinv};
        for (int i = 0; i < 7; ++i) if (outs[i]) cudaMemcpyAsync(outs[i], c.dense[i], sizeof(double) * nn, cudaMemcpyDeviceToHost, st);
        double* bouts[3] = {CinvBand, mphiBand, KinvBand};
        for (int i = 0; i < 3; ++i) if (bouts[i]) cudaMemcpyAsync(bouts[i], c.band[i], sizeof(double) * tab, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) rc = set_error(MAGI_ERR_CUDA, "D2H failed");
        if (repaired) { repaired[0] = c.rep_c.empty() ? 0 : c.rep_c[0]; repaired[1] = c.rep_k.empty() ? 0 : c.rep_k[0]; }
    }
    cudaStreamDestroy(st);
    cudaFree(pool);
    return rc;
}
