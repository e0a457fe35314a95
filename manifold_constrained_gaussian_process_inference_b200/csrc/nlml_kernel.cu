// GP hyper-parameter initialisation on device: the objective of
// Initialization.negative_log_marginal_likelihood (src/initialization.jl:72-176),
//   NLML(log var, log len, log sigma) = 0.5 * ( log|K_phi + (sigma^2 + jitter) I| + y^T (.)^-1 y + N log 2 pi )
// on the finite observations of one dimension, evaluated for a BATCH of candidate parameter vectors in one launch
// (one block per candidate: covariance build, in-place Cholesky with the pivot repair of cholesky(Positive, .),
// log-determinant and forward solve; the matrix lives in shared memory when it fits, else in a global work area).
// The Nelder-Mead driver (Optim.jl in the reference, src/initialization.jl:211-252) is host logic: initialization.py.
#include <cmath>
#include <cfloat>
#include "magi_internal.cuh"

namespace magi {

__global__ void __launch_bounds__(256) nlml_kernel(int kernel_id, int n, const double* __restrict__ t, const double* __restrict__ y,
                                                   double jitter, const double* __restrict__ log_params, double* __restrict__ out,
                                                   double* __restrict__ work, int use_smem) {
    extern __shared__ double sm[];
    __shared__ double red[8];
    __shared__ double s_piv;
    const int c = blockIdx.x;
    double* A = use_smem ? sm : work + (size_t)c * n * n;          // lower triangle, row-major A[i*n + j], j <= i
    double* z = use_smem ? sm + (size_t)n * n : work + (size_t)gridDim.x * n * n + (size_t)c * n;
    const double var = exp(log_params[3 * c]), len = exp(log_params[3 * c + 1]), sig = exp(log_params[3 * c + 2]);
    if (!isfinite(var) || !isfinite(len) || !isfinite(sig) || var <= 0 || len <= 0 || sig <= 0 || n <= 0) {   // :86-88
        if (threadIdx.x == 0) out[c] = INFINITY;
        return;
    }
    const double s = 1.0 / len, sqrt5 = sqrt(5.0), diag_add = sig * sig + jitter;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int i = idx / n, j = idx % n;
        if (j > i) continue;
        double k;
        if (kernel_id == MAGI_KERNEL_RBF) { const double d = t[i] * s - t[j] * s; k = var * exp(-(d * d) / 2.0); }
        else {
            const double r = fabs(t[i] * s - t[j] * s);
            if (kernel_id == MAGI_KERNEL_MATERN_NU12) k = var * exp(-r);                             // MaternKernel(nu), kernels.jl:109-118
            else if (kernel_id == MAGI_KERNEL_MATERN_NU32) k = var * ((1.0 + sqrt(3.0) * r) * exp(-sqrt(3.0) * r));
            else k = var * ((1.0 + sqrt5 * r + 5.0 * r * r / 3.0) * exp(-sqrt5 * r));
        }
        A[idx] = k + (i == j ? diag_add : 0.0);                                                  // :128
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) z[i] = y[i];
    __syncthreads();
    const double tol = (double)n * DBL_EPSILON * (var + diag_add);
    double logdet = 0.0;
    for (int j = 0; j < n; ++j) {
        if (threadIdx.x == 0) {
            double p = A[j * n + j];
            if (!(p > tol)) p = (fabs(p) > tol) ? fabs(p) : (tol > 0 ? tol : 1.0);              // cholesky(Positive, .) never throws
            s_piv = sqrt(p);
            A[j * n + j] = s_piv;
        }
        __syncthreads();
        const double r = s_piv;
        for (int i = j + 1 + threadIdx.x; i < n; i += blockDim.x) A[i * n + j] /= r;
        __syncthreads();
        const int m = n - j - 1;
        for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
            const int ii = j + 1 + idx / m, jj = j + 1 + idx % m;
            if (jj <= ii) A[ii * n + jj] -= A[ii * n + j] * A[jj * n + j];
        }
        __syncthreads();
        logdet += 2.0 * log(r);                                                                  // :138
    }
    // forward solve L z = y; y^T K^-1 y = ||z||^2   (:144-145)
    for (int i = 0; i < n; ++i) {
        double part = 0.0;
        for (int k = threadIdx.x; k < i; k += blockDim.x) part += A[i * n + k] * z[k];
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            double sum = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum += red[w];
            z[i] = (z[i] - sum) / A[i * n + i];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double quad = 0.0;
        for (int i = 0; i < n; ++i) quad += z[i] * z[i];
        const double v = 0.5 * (logdet + quad + n * log(2.0 * M_PI));                            // :150
        out[c] = isfinite(v) ? v : INFINITY;                                                     // :153-156
    }
}

}  // namespace magi

using namespace magi;

// NLML for n_cand candidate parameter vectors (log var, log len, log sigma) on the n finite observations (t, y) of one dimension.
extern "C" int magi_gp_nlml_batched(int kernel_id, int n, const double* t, const double* y, double jitter, int n_cand,
                                    const double* log_params, double* out, int device) {
    if (!t || !y || !log_params || !out || n < 1 || n_cand < 1) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_gp_nlml_batched: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_error(MAGI_ERR_CUDA, "magi_gp_nlml_batched: no CUDA device available (no CPU fallback)");
    if (cudaSetDevice(device) != cudaSuccess) return set_error(MAGI_ERR_CUDA, "cudaSetDevice failed");
    const size_t smem_need = sizeof(double) * ((size_t)n * n + n);
    const int use_smem = smem_need <= 200 * 1024;
    double* pool = nullptr;
    const size_t work = use_smem ? 0 : (size_t)n_cand * ((size_t)n * n + n);
    const size_t tot = 2 * (size_t)n + 4 * (size_t)n_cand + work;
    if (cudaMalloc(&pool, sizeof(double) * tot) != cudaSuccess) return set_error(MAGI_ERR_CUDA, "cudaMalloc failed");
    double *d_t = pool, *d_y = pool + n, *d_lp = pool + 2 * n, *d_out = d_lp + 3 * n_cand, *d_work = d_out + n_cand;
    cudaMemcpy(d_t, t, sizeof(double) * n, cudaMemcpyHostToDevice);
    cudaMemcpy(d_y, y, sizeof(double) * n, cudaMemcpyHostToDevice);
    cudaMemcpy(d_lp, log_params, sizeof(double) * 3 * n_cand, cudaMemcpyHostToDevice);
    static PerDeviceOnce once;
    if (once.need()) cudaFuncSetAttribute(nlml_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 64);
    nlml_kernel<<<n_cand, 256, use_smem ? smem_need : 0>>>(kernel_id, n, d_t, d_y, jitter, d_lp, d_out, d_work, use_smem);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, sizeof(double) * n_cand, cudaMemcpyDeviceToHost);
    cudaFree(pool);
    if (e != cudaSuccess) return cuda_error(e, "nlml_kernel");
    return MAGI_OK;
}
