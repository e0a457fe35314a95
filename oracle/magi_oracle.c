/* CPU oracle (C restatement) for the MAGI hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A loop-for-loop C restatement of the reference's per-leapfrog evaluation:
 *   log_likelihood_and_gradient_banded         /root/reference/src/likelihoods.jl:43-257
 *   LogDensityProblems.logdensity_and_gradient /root/reference/src/logdensityproblems_interface.jl:176-267
 * kept deliberately in the reference's structure (per-time-point ODE callback, four dgbmv-style band
 * products per dimension, Jacobians re-evaluated inside the dimension loop, scalar accumulation loops)
 * so that it is the honest stand-in for "the reference's CPU path" -- Julia is not installed here.
 * It is validated against oracle/magi_oracle.py (which is pinned to the reference's own test
 * known-answers) in tests/test_oracle_c.py, and timed by bench.py as cpu_baseline (kind "port").
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=native -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { MODEL_FN = 0, MODEL_HES1 = 1, MODEL_HES1LOG = 2, MODEL_HES1LOG_FIXG = 3, MODEL_HES1LOG_FIXF = 4,
       MODEL_HIV = 5, MODEL_PTRANS = 6, MODEL_LV = 7, MODEL_L96 = 8 };

#define MAXD 64
#define MAXK 16

/* ---- ODE models: src/ode_models.jl (f!(du,u,p,t), dfdx!(J,u,p,t) with J[i][j]=df_i/dx_j, dfdp -> D x k) ---- */
static void ode_f(int model, int D, const double *u, const double *p, double *du) {
    switch (model) {
    case MODEL_FN: { /* :39-47 */
        double V = u[0], R = u[1], a = p[0], b = p[1], c = p[2];
        du[0] = c * (V - (V * V * V) / 3.0 + R);
        du[1] = -1.0 / c * (V - a + b * R);
        break; }
    case MODEL_HES1: { /* :60-70 */
        double P = u[0], M = u[1], H = u[2];
        du[0] = -p[0] * P * H + p[1] * M - p[2] * P;
        du[1] = -p[3] * M + p[4] / (1 + P * P);
        du[2] = -p[0] * P * H + p[5] / (1 + P * P) - p[6] * H;
        break; }
    case MODEL_LV: { /* not in the reference */
        double x = u[0], y = u[1];
        du[0] = p[0] * x - p[1] * x * y;
        du[1] = p[2] * x * y - p[3] * y;
        break; }
    case MODEL_L96: {
        for (int i = 0; i < D; ++i) {
            int ip1 = (i + 1) % D, im1 = (i - 1 + D) % D, im2 = (i - 2 + D) % D;
            du[i] = (u[ip1] - u[im2]) * u[im1] - u[i] + p[0];
        }
        break; }
    default: for (int i = 0; i < D; ++i) du[i] = NAN;
    }
}

static void ode_dfdx(int model, int D, const double *u, const double *p, double *J /* D x D row-major */) {
    memset(J, 0, sizeof(double) * D * D);
    switch (model) {
    case MODEL_FN: { /* :248-262 */
        double V = u[0], b = p[1], c = p[2];
        J[0] = c * (1.0 - V * V); J[1] = c; J[2] = -1.0 / c; J[3] = -b / c;
        break; }
    case MODEL_HES1: { /* :312-336 */
        double P = u[0], H = u[2], opp = 1 + P * P;
        J[0] = -p[0] * H - p[2]; J[1] = p[1]; J[2] = -p[0] * P;
        J[3] = -p[4] * (2 * P) / (opp * opp); J[4] = -p[3]; J[5] = 0.0;
        J[6] = -p[0] * H - p[5] * (2 * P) / (opp * opp); J[7] = 0.0; J[8] = -p[0] * P - p[6];
        break; }
    case MODEL_LV: {
        double x = u[0], y = u[1];
        J[0] = p[0] - p[1] * y; J[1] = -p[1] * x; J[2] = p[2] * y; J[3] = p[2] * x - p[3];
        break; }
    case MODEL_L96: {
        for (int i = 0; i < D; ++i) {
            int ip1 = (i + 1) % D, im1 = (i - 1 + D) % D, im2 = (i - 2 + D) % D;
            J[i * D + ip1] += u[im1];
            J[i * D + im2] += -u[im1];
            J[i * D + im1] += u[ip1] - u[im2];
            J[i * D + i] += -1.0;
        }
        break; }
    default: for (int i = 0; i < D * D; ++i) J[i] = NAN;
    }
}

/* the reference's dfdp allocates and returns a fresh D x k matrix per call (ode_models.jl:280,355) */
static double *ode_dfdp(int model, int D, int k, const double *u, const double *p) {
    double *Jp = (double *)calloc((size_t)D * k, sizeof(double));
    switch (model) {
    case MODEL_FN: { /* :274-299 */
        double V = u[0], R = u[1], a = p[0], b = p[1], c = p[2];
        Jp[0 * 3 + 2] = V - (V * V * V) / 3.0 + R;
        Jp[1 * 3 + 0] = 1.0 / c;
        Jp[1 * 3 + 1] = -R / c;
        Jp[1 * 3 + 2] = (1.0 / (c * c)) * (V - a + b * R);
        break; }
    case MODEL_HES1: { /* :349-378 */
        double P = u[0], M = u[1], H = u[2], opp = 1 + P * P;
        Jp[0 * 7 + 0] = -P * H; Jp[0 * 7 + 1] = M; Jp[0 * 7 + 2] = -P;
        Jp[1 * 7 + 3] = -M; Jp[1 * 7 + 4] = 1.0 / opp;
        Jp[2 * 7 + 0] = -P * H; Jp[2 * 7 + 5] = 1.0 / opp; Jp[2 * 7 + 6] = -H;
        break; }
    case MODEL_LV: {
        double x = u[0], y = u[1];
        Jp[0 * 4 + 0] = x; Jp[0 * 4 + 1] = -x * y; Jp[1 * 4 + 2] = x * y; Jp[1 * 4 + 3] = -y;
        break; }
    case MODEL_L96: for (int i = 0; i < D; ++i) Jp[i] = 1.0; break;
    default: for (int i = 0; i < D * k; ++i) Jp[i] = NAN;
    }
    return Jp;
}

/* y = A x for a diagonal-major band table T[(b + j - i) * n + i] = A[i][j]; column-by-column axpy, the
 * order dgbmv('N') uses (BandedMatrices -> BLAS; call sites likelihoods.jl:129,132,133) */
static void gbmv_n(int n, int b, const double *T, const double *x, double *y) {
    for (int i = 0; i < n; ++i) y[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        double xj = x[j];
        int i0 = j - b < 0 ? 0 : j - b, i1 = j + b >= n ? n - 1 : j + b;
        for (int i = i0; i <= i1; ++i) y[i] += T[(size_t)(b + j - i) * n + i] * xj;
    }
}
/* y = A' x: dot product per column, the order dgbmv('T') uses (likelihoods.jl:192) */
static void gbmv_t(int n, int b, const double *T, const double *x, double *y) {
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        int i0 = j - b < 0 ? 0 : j - b, i1 = j + b >= n ? n - 1 : j + b;
        for (int i = i0; i <= i1; ++i) s += T[(size_t)(b + j - i) * n + i] * x[i];
        y[j] = s;
    }
}

/* likelihoods.jl:43-257.  X, Y are n x D column-major (time fastest).  bands: for each dim the three
 * (2b+1) x n diagonal-major tables mphiBand, KinvBand, CinvBand.  grad has n*D + k + D entries. */
int magi_oracle_loglik_grad(int n, int D, int k, int model, int b, const double *X, const double *theta,
                            const double *sigma, const double *Y, const double *mphiBand, const double *KinvBand,
                            const double *CinvBand, const double *beta, double *ll_out, double *grad) {
    if (D > MAXD || k > MAXK) return 1;
    size_t tab = (size_t)(2 * b + 1) * n;
    double ll = 0.0;
    int ngrad = n * D + k + D;
    double *fderiv = (double *)malloc(sizeof(double) * n * D);
    double *Ke_all = (double *)malloc(sizeof(double) * n * D);
    double *Cx_all = (double *)malloc(sizeof(double) * n * D);
    double *e0_all = (double *)malloc(sizeof(double) * n * D);
    unsigned char *fin_all = (unsigned char *)malloc((size_t)n * D);
    double *mx = (double *)malloc(sizeof(double) * n);
    double *e = (double *)malloc(sizeof(double) * n);
    double *mt = (double *)malloc(sizeof(double) * n);
    double u[MAXD], du[MAXD], sigma_sq[MAXD];
    double *J = (double *)malloc(sizeof(double) * D * D);
    memset(grad, 0, sizeof(double) * ngrad);
    for (int i = 0; i < n; ++i) { /* :89-95 */
        for (int d = 0; d < D; ++d) u[d] = X[(size_t)d * n + i];
        ode_f(model, D, u, theta, du);
        for (int d = 0; d < D; ++d) fderiv[(size_t)d * n + i] = du[d];
    }
    for (int d = 0; d < D; ++d) sigma_sq[d] = sigma[d] * sigma[d];
    for (int d = 0; d < D; ++d) { /* :111-152 */
        const double *xd = X + (size_t)d * n, *yd = Y + (size_t)d * n, *fd = fderiv + (size_t)d * n;
        double *e0 = e0_all + (size_t)d * n, *Ke = Ke_all + (size_t)d * n, *Cx = Cx_all + (size_t)d * n;
        unsigned char *fin = fin_all + (size_t)d * n;
        int nobs = 0;
        for (int i = 0; i < n; ++i) {
            fin[i] = isfinite(yd[i]) ? 1 : 0;
            e0[i] = fin[i] ? xd[i] - yd[i] : 0.0;
            nobs += fin[i];
        }
        gbmv_n(n, b, mphiBand + d * tab, xd, mx);
        for (int i = 0; i < n; ++i) e[i] = fd[i] - mx[i];
        gbmv_n(n, b, KinvBand + d * tab, e, Ke);
        gbmv_n(n, b, CinvBand + d * tab, xd, Cx);
        double sse = 0.0, eke = 0.0, xcx = 0.0;
        for (int i = 0; i < n; ++i) if (fin[i]) sse += e0[i] * e0[i];
        double ll_obs = -0.5 * sse / sigma_sq[d];
        if (nobs > 0) ll_obs -= 0.5 * nobs * log(2.0 * M_PI * sigma_sq[d]);
        ll += ll_obs / beta[2];
        for (int i = 0; i < n; ++i) eke += e[i] * Ke[i];
        ll += (-0.5 * eke) / beta[0];
        for (int i = 0; i < n; ++i) xcx += xd[i] * Cx[i];
        ll += (-0.5 * xcx) / beta[1];
    }
    double *gth = grad + (size_t)n * D, *gsig = grad + (size_t)n * D + k;
    for (int d = 0; d < D; ++d) { /* :168-247 */
        double *Ke = Ke_all + (size_t)d * n, *Cx = Cx_all + (size_t)d * n, *e0 = e0_all + (size_t)d * n;
        unsigned char *fin = fin_all + (size_t)d * n;
        double *gxd = grad + (size_t)d * n;
        for (int i = 0; i < n; ++i) if (fin[i]) gxd[i] -= (e0[i] / sigma_sq[d]) / beta[2];
        for (int i = 0; i < n; ++i) gxd[i] -= Cx[i] / beta[1];
        gbmv_t(n, b, mphiBand + d * tab, Ke, mt);
        for (int i = 0; i < n; ++i) gxd[i] += mt[i] / beta[0];
        for (int i = 0; i < n; ++i) { /* Jacobians re-evaluated per dimension, as the reference does (:199-209) */
            for (int j = 0; j < D; ++j) u[j] = X[(size_t)j * n + i];
            double kfe = Ke[i] / beta[0];
            ode_dfdx(model, D, u, theta, J);
            double *Jp = ode_dfdp(model, D, k, u, theta);
            for (int j = 0; j < D; ++j) grad[(size_t)j * n + i] -= J[d * D + j] * kfe;
            for (int q = 0; q < k; ++q) gth[q] -= Jp[d * k + q] * kfe;
            free(Jp);
        }
        if (sigma[d] > 0) { /* :229-246 */
            double sse = 0.0; int np = 0;
            for (int i = 0; i < n; ++i) if (fin[i]) { sse += e0[i] * e0[i]; np += 1; }
            if (np > 0) gsig[d] += (sse / sigma_sq[d] - np) / (sigma[d] * beta[2]);
        }
    }
    *ll_out = ll;
    free(fderiv); free(Ke_all); free(Cx_all); free(e0_all); free(fin_all); free(mx); free(e); free(mt); free(J);
    return 0;
}

/* logdensityproblems_interface.jl:176-267.  params = [vec(X); theta; log sigma] (P doubles). */
int magi_oracle_logdensity_and_gradient(int n, int D, int k, int model, int b, int sigma_is_fixed,
                                        const double *params, const double *sigma_init, const double *Y,
                                        const double *mphiBand, const double *KinvBand, const double *CinvBand,
                                        const double *beta, double *ll_out, double *grad_out) {
    int P = n * D + k + (sigma_is_fixed ? 0 : D);
    double sigma[MAXD], prior = 0.0;
    if (D > MAXD) return 1;
    if (sigma_is_fixed) {
        for (int d = 0; d < D; ++d) {
            sigma[d] = sigma_init[d];
            if (!isfinite(sigma[d]) || sigma[d] <= 0) { *ll_out = -INFINITY; for (int i = 0; i < P; ++i) grad_out[i] = NAN; return 0; }
        }
    } else {
        const double *ls = params + (size_t)n * D + k;
        for (int d = 0; d < D; ++d) {
            double v = ls[d]; v = v < -15.0 ? -15.0 : (v > 15.0 ? 15.0 : v);   /* :200 */
            sigma[d] = exp(v); prior += v;                      /* :201,206 */
        }
    }
    int ng = n * D + k + D;
    double *g = (double *)malloc(sizeof(double) * ng);
    double ll;
    magi_oracle_loglik_grad(n, D, k, model, b, params, params + (size_t)n * D, sigma, Y, mphiBand, KinvBand, CinvBand, beta, &ll, g);
    int ok = isfinite(ll);
    for (int i = 0; i < ng && ok; ++i) ok = isfinite(g[i]);
    if (!ok) { *ll_out = -INFINITY; for (int i = 0; i < P; ++i) grad_out[i] = 0.0; free(g); return 0; }   /* :222-226 */
    int nxt = n * D + k;
    for (int i = 0; i < nxt; ++i) grad_out[i] = g[i];
    double total = ll;
    if (!sigma_is_fixed) {
        total += prior;
        for (int d = 0; d < D; ++d) grad_out[nxt + d] = g[nxt + d] * sigma[d] + 1.0;   /* :249-253 */
    }
    ok = 1;
    for (int i = 0; i < P && ok; ++i) ok = isfinite(grad_out[i]);
    if (!ok) for (int i = 0; i < P; ++i) grad_out[i] = 0.0;                            /* :260-264 */
    *ll_out = total;
    free(g);
    return 0;
}

/* Batched driver: chains are independent, params is P x n_chains chain-contiguous. nthreads <= 0: all cores. */
int magi_oracle_batched(int n, int D, int k, int model, int b, int sigma_is_fixed, int n_chains, int nthreads,
                        const double *params, const double *sigma_init, const double *Y, const double *mphiBand,
                        const double *KinvBand, const double *CinvBand, const double *beta, double *ll, double *grad) {
    int P = n * D + k + (sigma_is_fixed ? 0 : D);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (int c = 0; c < n_chains; ++c)
        magi_oracle_logdensity_and_gradient(n, D, k, model, b, sigma_is_fixed, params + (size_t)c * P, sigma_init, Y,
                                            mphiBand, KinvBand, CinvBand, beta, ll + c, grad + (size_t)c * P);
    return 0;
}

int magi_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
