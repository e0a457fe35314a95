// K2: dense-mode evaluation.  Used when the band does not fit the windowed DMMA kernel (half-width > 32, in particular
// bandsize = n-1, which the reference treats as "dense": BandedMatrix == dense, test/test_gp.jl:551-585) and for models
// with many components (Lorenz-96).  Same mathematics as K1 (src/likelihoods.jl:43-257, interface.jl:176-267); the four
// products per dimension are (n x n) . (n x chains) GEMMs on DMMA tiles (gemm_f64.cu) that read the chain state through
// a strided view of the chain-contiguous parameter buffer, with the ODE / gradient work in two pointwise kernels:
//   MX = m~ X, CX = C~ X          (2 batched GEMMs over d)
//   E  = f(X, theta) - MX          (pointwise)
//   KE = K~ E                      (GEMM)
//   MT = m~^T KE                   (GEMM)
//   gradient, reductions, sigma transform and guards (one block per chain)
#include <cmath>
#include "magi_internal.cuh"
#include "gemm_f64.cuh"
#include "ode_models.cuh"
#include "dense_ode.cuh"

namespace magi {

#define DCK(call, what) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_error(e__, what); } while (0)

// band table (diagonal-major) -> dense n x n column-major with zeros outside the band (mat2band semantics)
__global__ void band_to_dense_kernel(const double* __restrict__ band, double* __restrict__ dense, int n, int b) {
    const int d = blockIdx.z;
    const size_t nn = (size_t)n * n, tab = (size_t)(2 * b + 1) * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < nn; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n), j = (int)(idx / n);
        const int off = j - i;
        dense[d * nn + idx] = (off >= -b && off <= b) ? band[d * tab + (size_t)(b + off) * n + i] : 0.0;
    }
}

template <int MODEL>
__global__ void dense_e_kernel(const double* __restrict__ params, long long pitch, int n, int D, int n_chains,
                               const double* MX, double* E) {   // E may alias MX (in place)
    constexpr int K = DenseOde<MODEL>::K;
    const int c = blockIdx.x;                    // chains on grid.x: no 65535 limit
    const double* xp = params + (size_t)c * pitch;
    double th[DenseOde<MODEL>::KX];
#pragma unroll
    for (int i = 0; i < K; ++i) th[i] = xp[(size_t)n * D + i];
    DenseOde<MODEL>::prepare(th);
    const size_t plane = (size_t)n * n_chains;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        auto x = [&](int dd) { return xp[(size_t)dd * n + i]; };
        for (int d = 0; d < D; ++d) {
            const size_t o = (size_t)d * plane + (size_t)c * n + i;
            E[o] = DenseOde<MODEL>::f(d, x, th, D) - MX[o];        // likelihoods.jl:130
        }
    }
}

struct DenseGradArgs {
    int n, D, K, P, n_chains, sigma_is_fixed, sigma_invalid;
    long long pitch;
    const double* params; double* ll; double* grad;
    const double *E, *KE, *CX, *MT;
    const double* yobs; const int* nobs; const double* sigma_init;
    double beta[3], inv_beta[3];
    double* part;       // [n_chains][D][4 + K]: e.Ke, x.Cx, sse, bad flag, theta-gradient partials of one (chain, dimension)
};

// Gradient with respect to the states of ONE (chain, dimension) and that dimension's partial sums (likelihoods.jl:139-221);
// grid (n_chains, D): a model with 64 components gives 64 blocks per chain instead of one block looping over the dimensions.
template <int MODEL>
__global__ void __launch_bounds__(256) dense_grad_part_kernel(const DenseGradArgs a) {
    constexpr int K = DenseOde<MODEL>::K;
    __shared__ double sh[8][4 + K];
    const int c = blockIdx.x, d = blockIdx.y, n = a.n, D = a.D;
    if (a.sigma_invalid) return;                                      // (the finalize kernel writes -Inf / NaN, interface.jl:192-195)
    const double* xp = a.params + (size_t)c * a.pitch;
    double* gp = a.grad ? a.grad + (size_t)c * a.pitch : nullptr;
    const size_t plane = (size_t)n * a.n_chains, base = (size_t)c * n;
    const int nxt = n * D + K;
    double th[DenseOde<MODEL>::KX];
#pragma unroll
    for (int i = 0; i < K; ++i) th[i] = xp[(size_t)n * D + i];
    DenseOde<MODEL>::prepare(th);
    const double inv_b1 = a.inv_beta[0], inv_b2 = a.inv_beta[1], inv_b3 = a.inv_beta[2];
    double gth[K];
#pragma unroll
    for (int i = 0; i < K; ++i) gth[i] = 0.0;
    double s;
    if (a.sigma_is_fixed) s = a.sigma_init[d];
    else {
        const double raw = xp[nxt + d];
        s = isnan(raw) ? raw : exp(fmin(fmax(raw, -15.0), 15.0));     // interface.jl:200
    }
    const double inv_sig2 = 1.0 / (s * s);
    const double* Ed = a.E + (size_t)d * plane + base;
    const double* KEd = a.KE + (size_t)d * plane + base;
    const double* CXd = a.CX + (size_t)d * plane + base;
    const double* MTd = a.MT + (size_t)d * plane + base;
    const double* yd = a.yobs + (size_t)d * n;
    double eke = 0.0, xcx = 0.0, sse = 0.0;
    bool bad = false;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        auto x = [&](int dd) { return xp[(size_t)dd * n + i]; };
        auto w = [&](int dd) { return a.KE[(size_t)dd * plane + base + i] * inv_b1; };   // likelihoods.jl:201
        const double xdv = xp[(size_t)d * n + i], y = yd[i], cx = CXd[i], ke = KEd[i];
        const bool fin = isfinite(y);
        const double e0 = fin ? xdv - y : 0.0;
        double gv = 0.0;
        if (fin) gv -= (e0 * inv_sig2) * inv_b3;                   // likelihoods.jl:179
        gv -= cx * inv_b2;                                         // :186
        gv += MTd[i] * inv_b1;                                     // :194
        DenseOde<MODEL>::jx_col_sub(d, x, w, th, D, gv);           // :214-216
        DenseOde<MODEL>::jth_row_sub(d, x, th, D, ke * inv_b1, gth);   // :219-221
        eke += Ed[i] * ke;
        xcx += xdv * cx;
        sse += e0 * e0;
        bad |= gp && !isfinite(gv);            // value-only calls look at the log density alone (interface.jl:155-160)
        if (gp) gp[(size_t)d * n + i] = gv;
    }
    // one hand-over for all 4 + K sums: warp butterflies, one row of partials per warp, then thread j adds column j
    double v[4 + K];
    v[0] = eke; v[1] = xcx; v[2] = sse; v[3] = bad ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) v[4 + i] = gth[i];
#pragma unroll
    for (int j = 0; j < 4 + K; ++j)
        for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    const int wp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int j = 0; j < 4 + K; ++j) sh[wp][j] = v[j];
    }
    __syncthreads();
    if (threadIdx.x < 4 + K) {
        double r = 0.0;
        for (int i = 0; i < 8; ++i) r += sh[i][threadIdx.x];
        a.part[((size_t)c * D + d) * (4 + K) + threadIdx.x] = r;
    }
}

// ---- Lorenz-96 (D = 64) versions of the two pointwise stages.  The generic kernels above read every state value 4-5 times (one
// block per (chain, dimension): 15 loads per point); a Lorenz-96 component couples only to its four neighbours, so
//   * E: one thread per (chain, time) walks the dimensions with a rolling 4-value register window: one state load per point;
//   * gradient: a block takes kL96Group consecutive dimensions of a chain and loads their kL96Group + 4 neighbouring state and KE
//     values ONCE per time point; same thread <-> time mapping, same operation order and the same reduction tree as the generic
//     kernel, hence the same bits. ----
constexpr int kL96Group = 4;

__global__ void __launch_bounds__(256) lorenz96_e_kernel(const double* __restrict__ params, long long pitch, int n, int D, int n_chains,
                                                         const double* MX, double* E) {   // E may alias MX (in place)
    const int c = blockIdx.x;
    const double* xp = params + (size_t)c * pitch;
    const double F = xp[(size_t)n * D];
    const size_t plane = (size_t)n * n_chains;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        const double* xi = xp + i;
        double xm2 = xi[(size_t)(D - 2) * n], xm1 = xi[(size_t)(D - 1) * n], x0 = xi[0], xp1 = xi[(size_t)(1 % D) * n];
        for (int d = 0; d < D; ++d) {
            const size_t o = (size_t)d * plane + (size_t)c * n + i;
            E[o] = ((xp1 - xm2) * xm1 - x0 + F) - MX[o];             // DenseOde<L96>::f, likelihoods.jl:130
            xm2 = xm1; xm1 = x0; x0 = xp1;
            xp1 = xi[(size_t)((d + 2) % D) * n];
        }
    }
}

__global__ void __launch_bounds__(256, 2) lorenz96_grad_group_kernel(const DenseGradArgs a) {
    constexpr int G = kL96Group, NV = 5;                             // e.Ke, x.Cx, sse, bad flag, dF
    __shared__ double sh[8][G * NV];
    __shared__ double s_inv_sig2[G];
    __shared__ int s_dim[G + 4];                                     // dimensions d0 - 2 .. d0 + G + 1 (cyclic)
    const int c = blockIdx.x, d0 = blockIdx.y * G, n = a.n, D = a.D;
    if (a.sigma_invalid) return;
    const double* xp = a.params + (size_t)c * a.pitch;
    double* gp = a.grad ? a.grad + (size_t)c * a.pitch : nullptr;
    const size_t plane = (size_t)n * a.n_chains, base = (size_t)c * n;
    const int nxt = n * D + 1;
    const double inv_b1 = a.inv_beta[0], inv_b2 = a.inv_beta[1], inv_b3 = a.inv_beta[2];
    if (threadIdx.x < G + 4) s_dim[threadIdx.x] = ((d0 - 2 + (int)threadIdx.x) % D + D) % D;
    if (threadIdx.x < G) {
        const int d = d0 + threadIdx.x < D ? d0 + threadIdx.x : D - 1;
        double s;
        if (a.sigma_is_fixed) s = a.sigma_init[d];
        else {
            const double raw = xp[nxt + d];
            s = isnan(raw) ? raw : exp(fmin(fmax(raw, -15.0), 15.0));     // interface.jl:200
        }
        s_inv_sig2[threadIdx.x] = 1.0 / (s * s);
    }
    __syncthreads();
    double acc[G][NV - 1];
    unsigned badmask = 0;
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int j = 0; j < NV - 1; ++j) acc[g][j] = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double xr[G + 4], ker[G + 4];
#pragma unroll
        for (int k = 0; k < G + 4; ++k) {
            const int dd = s_dim[k];
            xr[k] = xp[(size_t)dd * n + i];
            ker[k] = a.KE[(size_t)dd * plane + base + i];
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int d = d0 + g;
            if (d < D) {
                const size_t o = (size_t)d * plane + base + i;
                const double xdv = xr[g + 2], y = a.yobs[(size_t)d * n + i], cx = a.CX[o], ke = ker[g + 2];
                const bool fin = isfinite(y);
                const double e0 = fin ? xdv - y : 0.0;
                double gv = 0.0;
                if (fin) gv -= (e0 * s_inv_sig2[g]) * inv_b3;        // likelihoods.jl:179
                gv -= cx * inv_b2;                                   // :186
                gv += a.MT[o] * inv_b1;                              // :194
                gv -= xr[g] * (ker[g + 1] * inv_b1);                 // :214-216 with w = Ke / beta1 (:201); column d of the Jacobian: d f_{d-1} / d x_d = x_{d-2}
                gv -= (-1.0) * (ker[g + 2] * inv_b1);                //   d f_d / d x_d = -1
                gv -= (xr[g + 4] - xr[g + 1]) * (ker[g + 3] * inv_b1);   //   d f_{d+1} / d x_d = x_{d+2} - x_{d-1}
                gv -= (-xr[g + 3]) * (ker[g + 4] * inv_b1);          //   d f_{d+2} / d x_d = -x_{d+1}
                acc[g][3] -= ke * inv_b1;                            // :219-221 (d f_d / d F = 1)
                acc[g][0] += a.E[o] * ke;
                acc[g][1] += xdv * cx;
                acc[g][2] += e0 * e0;
                if (gp && !isfinite(gv)) badmask |= 1u << g;
                if (gp) gp[(size_t)d * n + i] = gv;
            }
        }
    }
    const int wp = threadIdx.x >> 5;
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            // part layout of the generic kernel: e.Ke, x.Cx, sse, bad flag, theta-gradient partial
            double v = j < 3 ? acc[g][j] : (j == 3 ? (((badmask >> g) & 1u) ? 1.0 : 0.0) : acc[g][3]);
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) sh[wp][g * NV + j] = v;
        }
    __syncthreads();
    if (threadIdx.x < G * NV) {
        const int g = threadIdx.x / NV, j = threadIdx.x % NV;
        if (d0 + g < D) {
            double r = 0.0;
            for (int i = 0; i < 8; ++i) r += sh[i][threadIdx.x];
            a.part[((size_t)c * D + d0 + g) * NV + j] = r;
        }
    }
}

// Per chain: the log density in the reference's order of accumulation over the dimensions, sigma gradient, log-sigma
// transform, guards (interface.jl:192-264); the per-dimension terms (log, divisions) are formed by D threads in parallel.
template <int MODEL>
__global__ void __launch_bounds__(256) dense_finalize_kernel(const DenseGradArgs a) {
    constexpr int K = DenseOde<MODEL>::K;
    extern __shared__ double fin[];                                   // [D][4]: ll_obs / beta3, -eke / 2 beta1, -xcx / 2 beta2, d/d log sigma
    __shared__ int sflag;
    const int c = blockIdx.x, n = a.n, D = a.D;
    const double* xp = a.params + (size_t)c * a.pitch;
    double* gp = a.grad ? a.grad + (size_t)c * a.pitch : nullptr;
    const int nxt = n * D + K, P = a.P;
    if (a.sigma_invalid) {                                            // interface.jl:192-195
        if (threadIdx.x == 0) a.ll[c] = -INFINITY;
        if (gp) for (int i = threadIdx.x; i < P; i += blockDim.x) gp[i] = NAN;
        return;
    }
    if (threadIdx.x == 0) sflag = 0;
    __syncthreads();
    int flag = 0;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const double* r = a.part + ((size_t)c * D + d) * (4 + K);
        double s;
        if (a.sigma_is_fixed) s = a.sigma_init[d];
        else {
            const double raw = xp[nxt + d];
            s = isnan(raw) ? raw : exp(fmin(fmax(raw, -15.0), 15.0));
        }
        const double s2 = s * s, sse = r[2];
        const int nobs = a.nobs[d];
        double ll_obs = -0.5 * sse / s2;                              // likelihoods.jl:139
        if (nobs > 0) ll_obs -= 0.5 * nobs * log(2.0 * M_PI * s2);   // :141
        fin[4 * d + 0] = ll_obs / a.beta[2];                          // :143
        fin[4 * d + 1] = (-0.5 * r[0]) / a.beta[0];                   // :146-147
        fin[4 * d + 2] = (-0.5 * r[1]) / a.beta[1];                   // :150-151
        const double gsig = (s > 0 && nobs > 0) ? (sse / s2 - nobs) / (s * a.beta[2]) : 0.0;   // :229-246
        const double gls = gsig * s + 1.0;                            // interface.jl:249-253
        fin[4 * d + 3] = gls;
        if (r[3] != 0.0 || (gp && !isfinite(gsig))) flag |= 1;
        if (!a.sigma_is_fixed && !isfinite(gls)) flag |= 2;
    }
    if (flag) atomicOr(&sflag, flag);
    __syncthreads();
    __shared__ double s_ll;
    static_assert(K <= 16, "theta gradient staging");
    __shared__ double s_gth[16];
    if (threadIdx.x == 0) {
        double ll = 0.0, prior = 0.0, gth[K];
#pragma unroll
        for (int i = 0; i < K; ++i) gth[i] = 0.0;
        for (int d = 0; d < D; ++d) {                                 // the reference's order of accumulation
            ll += fin[4 * d + 0]; ll += fin[4 * d + 1]; ll += fin[4 * d + 2];
            const double* r = a.part + ((size_t)c * D + d) * (4 + K);
#pragma unroll
            for (int i = 0; i < K; ++i) gth[i] += r[4 + i];
            if (!a.sigma_is_fixed) {
                const double raw = xp[nxt + d];
                prior += isnan(raw) ? raw : fmin(fmax(raw, -15.0), 15.0);   // interface.jl:206
            }
        }
        bool bad = !isfinite(ll);
        if (gp) {
#pragma unroll
            for (int i = 0; i < K; ++i) bad |= !isfinite(gth[i]);
        }
        if (bad) atomicOr(&sflag, 1);
        s_ll = a.sigma_is_fixed ? ll : ll + prior;
#pragma unroll
        for (int i = 0; i < K; ++i) s_gth[i] = gth[i];
    }
    __syncthreads();
    const int fl = sflag;
    if (fl & 1) {                                                      // interface.jl:222-226
        if (threadIdx.x == 0) a.ll[c] = -INFINITY;
        if (gp) for (int i = threadIdx.x; i < P; i += blockDim.x) gp[i] = 0.0;
        return;
    }
    if (threadIdx.x == 0) a.ll[c] = s_ll;
    if (gp) {
        if (fl & 2) { for (int i = threadIdx.x; i < P; i += blockDim.x) gp[i] = 0.0; }   // interface.jl:260-264
        else {
            for (int i = threadIdx.x; i < K; i += blockDim.x) gp[n * D + i] = s_gth[i];
            if (!a.sigma_is_fixed) for (int d = threadIdx.x; d < D; d += blockDim.x) gp[nxt + d] = fin[4 * d + 3];
        }
    }
}

template <int MODEL>
static int dense_pointwise(magi_handle* h, int n_chains, const double* params, long long pitch, double* ll, double* grad,
                           const double* MX, double* E, const double* KE, const double* CX, const double* MT, int stage, cudaStream_t st) {
    if (stage == 0) {
        dim3 grid(n_chains, (h->n + 255) / 256);
        if constexpr (MODEL == MAGI_MODEL_L96) lorenz96_e_kernel<<<grid, 256, 0, st>>>(params, pitch, h->n, h->D, n_chains, MX, E);
        else dense_e_kernel<MODEL><<<grid, 256, 0, st>>>(params, pitch, h->n, h->D, n_chains, MX, E);
    } else {
        DenseGradArgs a;
        a.n = h->n; a.D = h->D; a.K = h->K; a.P = h->P; a.n_chains = n_chains; a.sigma_is_fixed = h->sigma_is_fixed; a.sigma_invalid = h->sigma_invalid;
        a.pitch = pitch; a.params = params; a.ll = ll; a.grad = grad; a.E = E; a.KE = KE; a.CX = CX; a.MT = MT;
        a.yobs = h->d_yobs; a.nobs = h->d_nobs; a.sigma_init = h->d_sigma_init;
        for (int i = 0; i < 3; ++i) { a.beta[i] = h->beta[i]; a.inv_beta[i] = 1.0 / h->beta[i]; }
        const size_t need = (size_t)n_chains * h->D * (4 + h->K);
        if (need > h->dense_part_cap) {
            if (h->d_dense_part) cudaFree(h->d_dense_part);
            h->d_dense_part = nullptr; h->dense_part_cap = 0;
            DCK(cudaMalloc(&h->d_dense_part, sizeof(double) * need), "cudaMalloc per-dimension partial sums");
            h->dense_part_cap = need;
        }
        a.part = h->d_dense_part;
        if constexpr (MODEL == MAGI_MODEL_L96) lorenz96_grad_group_kernel<<<dim3(n_chains, (h->D + kL96Group - 1) / kL96Group), 256, 0, st>>>(a);
        else dense_grad_part_kernel<MODEL><<<dim3(n_chains, h->D), 256, 0, st>>>(a);
        dense_finalize_kernel<MODEL><<<n_chains, 256, sizeof(double) * 4 * h->D, st>>>(a);
        h->launches++;
    }
    DCK(cudaGetLastError(), "dense pointwise kernel");
    h->launches++;
    return MAGI_OK;
}

static int dense_pointwise_dispatch(magi_handle* h, int n_chains, const double* params, long long pitch, double* ll, double* grad,
                                    const double* MX, double* E, const double* KE, const double* CX, const double* MT, int stage, cudaStream_t st) {
    switch (h->model) {
#define MAGI_CASE(M) case M: return dense_pointwise<M>(h, n_chains, params, pitch, ll, grad, MX, E, KE, CX, MT, stage, st);
    MAGI_CASE(MAGI_MODEL_FN) MAGI_CASE(MAGI_MODEL_HES1) MAGI_CASE(MAGI_MODEL_LV) MAGI_CASE(MAGI_MODEL_L96)
    MAGI_CASE(MAGI_MODEL_HES1LOG) MAGI_CASE(MAGI_MODEL_HES1LOG_FIXG) MAGI_CASE(MAGI_MODEL_HES1LOG_FIXF) MAGI_CASE(MAGI_MODEL_HIV) MAGI_CASE(MAGI_MODEL_PTRANS)
#undef MAGI_CASE
    default: return set_error(MAGI_ERR_UNSUPPORTED, "dense mode: unknown model");
    }
}

// Many-component models with a band the DMMA tiling covers (half-width <= 32): the four products per dimension as band
// products on the fragment tables (band_product.cu) instead of band-truncated GEMMs; same pointwise stages.
static int eval_band_products_dev(magi_handle* h, int n_chains, const double* params, long long pitch, double* ll, double* grad, cudaStream_t st) {
    const int n = h->n, D = h->D;
    const size_t plane = (size_t)n * n_chains;
    if (!h->d_fragtab_bp) { DCK(cudaMalloc(&h->d_fragtab_bp, sizeof(double) * fragtab_doubles(n, h->b, D)), "cudaMalloc fragment tables"); h->frag_bp_dirty = true; }
    if (h->frag_bp_dirty) {
        DCK(launch_build_fragtab(h->d_band[0], h->d_band[1], h->d_band[2], h->d_fragtab_bp, n, h->b, D, true, 1.0, 1.0, st), "build_fragtab");
        h->launches++; h->frag_bp_dirty = false;
    }
    const size_t need = 4 * plane * D;
    if (need > h->dense_work_cap) {
        if (h->d_dense_work) cudaFree(h->d_dense_work);
        h->d_dense_work = nullptr; h->dense_work_cap = 0;
        DCK(cudaMalloc(&h->d_dense_work, sizeof(double) * need), "cudaMalloc work space");
        h->dense_work_cap = need;
    }
    double* MXE = h->d_dense_work;            // MX, then E in place
    double* KE = MXE + plane * D;
    double* CX = KE + plane * D;
    double* MT = CX + plane * D;
    auto bp = [&](int view, const double* in, long long cs, long long ds, double* out) {
        cudaError_t e = launch_band_product(h->d_fragtab_bp, view, in, cs, ds, out, (long long)plane, n, h->b, D, n_chains, h->sm_count, st);
        h->launches++;
        return e;
    };
    DCK(bp(0, params, pitch, n, MXE), "band product m~ X");                      // likelihoods.jl:129
    DCK(bp(1, params, pitch, n, CX), "band product C~ X");                       // :133
    int rc = dense_pointwise_dispatch(h, n_chains, params, pitch, ll, grad, MXE, MXE, nullptr, nullptr, nullptr, 0, st);   // E = f - MX (:130)
    if (rc) return rc;
    DCK(bp(2, MXE, n, (long long)plane, KE), "band product K~ E");               // :132
    DCK(bp(3, KE, n, (long long)plane, MT), "band product m~^T KE");             // :192
    return dense_pointwise_dispatch(h, n_chains, params, pitch, ll, grad, nullptr, MXE, KE, CX, MT, 1, st);
}

int eval_dense_dev(magi_handle* h, int n_chains, const double* params, long long pitch, double* ll, double* grad, cudaStream_t st) {
    if (h->geom.HB <= kMaxHB && h->b < h->n - 1) return eval_band_products_dev(h, n_chains, params, pitch, ll, grad, st);
    const int n = h->n, D = h->D;
    const size_t nn = (size_t)n * n, plane = (size_t)n * n_chains;
    // dense (band-truncated) operators, rebuilt when the band tables change
    if (!h->d_dense_ops) {
        DCK(cudaMalloc(&h->d_dense_ops, sizeof(double) * 3 * nn * D), "cudaMalloc dense operators");
        h->dense_band_dirty = true;
    }
    if (h->dense_band_dirty) {
        size_t blocks = (nn + 255) / 256; if (blocks > 148 * 16) blocks = 148 * 16;
        for (int t = 0; t < 3; ++t) {
            band_to_dense_kernel<<<dim3((unsigned)blocks, 1, D), 256, 0, st>>>(h->d_band[t], h->d_dense_ops + (size_t)t * nn * D, n, h->b);
            h->launches++;
        }
        DCK(cudaGetLastError(), "band_to_dense_kernel");
        h->dense_band_dirty = false;
    }
    const size_t need = 4 * plane * D;
    if (need > h->dense_work_cap) {
        if (h->d_dense_work) cudaFree(h->d_dense_work);
        h->d_dense_work = nullptr; h->dense_work_cap = 0;
        DCK(cudaMalloc(&h->d_dense_work, sizeof(double) * need), "cudaMalloc dense work space");
        h->dense_work_cap = need;
    }
    if (!h->d_sk_work) {
        DCK(cudaMalloc(&h->d_sk_work, sizeof(double) * kStreamKWorkDoubles), "cudaMalloc stream-K work space");
        DCK(cudaMalloc(&h->d_sk_flags, sizeof(unsigned) * kStreamKSlots), "cudaMalloc stream-K flags");
        DCK(cudaMemsetAsync(h->d_sk_flags, 0, sizeof(unsigned) * kStreamKSlots, st), "memset stream-K flags");
    }
    double* MXE = h->d_dense_work;            // MX, then E in place
    double* KE = MXE + plane * D;
    double* CX = KE + plane * D;
    double* MT = CX + plane * D;
    const double* Cinv = h->d_dense_ops;                  // [D][n x n]
    const double* Mphi = h->d_dense_ops + nn * D;
    const double* Kinv = h->d_dense_ops + 2 * nn * D;
    auto gemm = [&](const double* A, bool tA, const double* B, long long rsB, long long csB, long long bsB, double* C) {
        GemmArgs g{};
        g.A = A; g.rsA = tA ? n : 1; g.csA = tA ? 1 : n; g.bsA1 = (long long)nn; g.bsA2 = 0;
        g.B = B; g.rsB = rsB; g.csB = csB; g.bsB1 = bsB; g.bsB2 = 0;
        g.C = C; g.rsC = 1; g.csC = n; g.bsC1 = (long long)plane; g.bsC2 = 0;
        g.M = n; g.N = n_chains; g.K = n; g.nb1 = 1 << 30; g.alpha = 1.0; g.beta = 0.0;
        g.a_band = (h->b < n - 1) ? h->b : 0;       // band-truncated operators: only the k-tiles that meet the band are multiplied
        g.sk_work = h->d_sk_work; g.sk_flags = h->d_sk_flags; g.sk_epoch = ++h->sk_epoch; g.extra_launches = &h->launches;
        if (g.sk_epoch == 0) g.sk_epoch = ++h->sk_epoch;
        cudaError_t e = launch_gemm(g, D, st);
        h->launches++;
        return e;
    };
    // X_d(i, c) = params[c * pitch + d * n + i]
    {   // MX = m~ X and CX = C~ X in ONE launch (2 D batched products sharing the B operand X: batch z = d + D * which):
        // 4 D instead of 2 x 2 D tile rows for the persistent stream-K grid to cut evenly
        GemmArgs g{};
        g.A = Cinv; g.rsA = 1; g.csA = n; g.bsA1 = (long long)nn; g.bsA2 = (long long)(Mphi - Cinv);
        g.B = params; g.rsB = 1; g.csB = pitch; g.bsB1 = n; g.bsB2 = 0;
        g.C = CX; g.rsC = 1; g.csC = n; g.bsC1 = (long long)plane; g.bsC2 = (long long)(MXE - CX);
        g.M = n; g.N = n_chains; g.K = n; g.nb1 = D; g.alpha = 1.0; g.beta = 0.0;
        g.a_band = (h->b < n - 1) ? h->b : 0;
        g.sk_work = h->d_sk_work; g.sk_flags = h->d_sk_flags; g.sk_epoch = ++h->sk_epoch; g.extra_launches = &h->launches;
        if (g.sk_epoch == 0) g.sk_epoch = ++h->sk_epoch;
        DCK(launch_gemm(g, 2 * D, st), "dense gemm [C~ | m~] X");
        h->launches++;
    }
    int rc = dense_pointwise_dispatch(h, n_chains, params, pitch, ll, grad, MXE, MXE, nullptr, nullptr, nullptr, 0, st);
    if (rc) return rc;
    DCK(gemm(Kinv, false, MXE, 1, n, (long long)plane, KE), "dense gemm K~ E");
    DCK(gemm(Mphi, true, KE, 1, n, (long long)plane, MT), "dense gemm m~^T KE");
    return dense_pointwise_dispatch(h, n_chains, params, pitch, ll, grad, nullptr, MXE, KE, CX, MT, 1, st);
}

}  // namespace magi
