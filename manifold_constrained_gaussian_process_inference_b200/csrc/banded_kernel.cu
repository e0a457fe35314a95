// K1: fused banded log-posterior + gradient on FP64 tensor cores (DMMA.8x8x4), and K6: band -> fragment tables.
//
// Replaces, for a batch of independent chains, the reference's
//   log_likelihood_and_gradient_banded          src/likelihoods.jl:43-257
//   LogDensityProblems.logdensity_and_gradient  src/logdensityproblems_interface.jl:176-267
//
// Formulation.  For every dimension d the four band products (m~ x, K~ e, C~ x, m~^T Ke; likelihoods.jl:129,132,133,192)
// are written as (chains x time) = (chains x time) . (band table) products: 8 chains form the M extent of a
// DMMA.8x8x4, 8 output times its N extent, and the contraction runs over 4-time chunks.  A warp owns one
// (chain-group, dimension) task and sweeps the time axis once per phase; the operand (x, e, Ke) lives in a register
// window of WN chunks that slides by one 8-time tile per step, so every state value is read from memory once per
// sweep and the only per-DMMA load is the 256-byte table fragment (shared by every chain on the GPU: L1/L2 hits).
// The time->slot permutation (lane (gid,q) owns times 8J+q and 8J+q+4 of tile J) makes the C fragment of one
// product directly usable as the A fragment of the next, so x -> e -> Ke never leaves registers.
//
//   phase A1 (per task):  mx = m~ x_d;  e = f_d(x, theta) - mx;  Ke = K~ e  -> Ke to scratch; sum e.Ke
//   phase A2 (per task):  Cx = C~ x_d;  mt = m~^T Ke_d;  pointwise gradient incl. the ODE Jacobian terms, which need
//                         Ke of ALL dimensions at the same time point (hence the block-wide barrier in between)
//   final   (per chain):  log-likelihood assembly in the reference's term order, sigma gradient, log-sigma
//                         transform and the per-chain -Inf / zero-gradient guards.
#include <cmath>
#include <cstdlib>
#include "magi_common.cuh"
#include "ode_models.cuh"

namespace magi {

// ------------------------------------------------------------------------------------------------------------
// K6: fragment tables.  fragtab[view][d][J][hh][lane], lane = 4*gid + q holds the B-operand entry
//   T[in = 4*(2J - HB + hh) + q][out = 8J + (gid>>1) + 4*(gid&1)]
// with T[in][out] = A[out][in] for y = A x (views 0: m~, 1: C~, 2: K~) and T[in][out] = m~[in][out] for view 3 (m~^T).
// Band rule |in - out| <= b as mat2band (gaussian_process.jl:70-74, 358-360).  Input tables are diagonal-major.
// ------------------------------------------------------------------------------------------------------------
__global__ void build_fragtab_kernel(const double* __restrict__ band_cinv, const double* __restrict__ band_mphi,
                                     const double* __restrict__ band_kinv, double* __restrict__ fragtab,
                                     int n, int b, int D, int HB, int NCH, int NT) {
    const size_t per_view = (size_t)D * NT * NCH * 32;
    const size_t total = 4 * per_view;
    const size_t tab = (size_t)(2 * b + 1) * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int lane = (int)(idx & 31);
        size_t r = idx >> 5;
        int hh = (int)(r % NCH); r /= NCH;
        int J = (int)(r % NT); r /= NT;
        int d = (int)(r % D);
        int view = (int)(r / D);
        int gid = lane >> 2, q = lane & 3;
        int o = 8 * J + (gid >> 1) + 4 * (gid & 1);
        int i = 4 * (2 * J - HB + hh) + q;
        double v = 0.0;
        if (o < n && i >= 0 && i < n && abs(i - o) <= b) {
            const double* src = (view == 1) ? band_cinv : (view == 2 ? band_kinv : band_mphi);
            src += (size_t)d * tab;
            v = (view == 3) ? src[(size_t)(b + (o - i)) * n + i] : src[(size_t)(b + (i - o)) * n + o];
        }
        fragtab[idx] = v;
    }
}

cudaError_t launch_build_fragtab(const double* band_cinv, const double* band_mphi, const double* band_kinv, double* fragtab,
                                 int n, int b, int D, cudaStream_t st) {
    BandGeom g = band_geom(n, b);
    size_t total = (size_t)4 * D * g.NT * g.NCH * 32;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    build_fragtab_kernel<<<blocks, 256, 0, st>>>(band_cinv, band_mphi, band_kinv, fragtab, n, b, D, g.HB, g.NCH, g.NT);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double quad_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// One block = G chain-groups (8 chains each) x DW dimension slots x H time segments, one warp per (group, dim, segment)
// task.  Three barrier-separated phases; e and Ke live in a scratch area laid out in fragment order
// scr[(g*D + d)*NT + tile][2][32] (shared memory when it fits, else global/L2), so that
//   * the time axis splits across warps with no halo recomputation (a neighbour's e / Ke is read from the scratch),
//   * the A operands of K~ e and m~^T Ke are plain conflict-free LDS, with no register window,
//   * the register budget stays <= 128, i.e. 16 resident warps per SM (the kernel is latency-, not issue-bound).
//   P1: e  = f_d(x, theta) - m~ x_d          (x_d in a sliding register window fed from global memory)
//   P2: Ke = K~ e ;  sum e.Ke
//   P3: Cx = C~ x_d ; mt = m~^T Ke_d ; pointwise gradient incl. the ODE Jacobian terms (need Ke of all dimensions)
template <int MODEL, int HB>
__global__ void __launch_bounds__(256, 2) banded_logpost_kernel(const BandedArgs a) {
    using M = Ode<MODEL>;
    constexpr int D = M::D, K = M::K;
    constexpr int NCH = 2 * HB + 2, LAGT = (HB + 1) / 2, WN = 2 * LAGT + 2 + HB, CH2 = (HB + 1) / 2;
    constexpr int RED = 4 + K;   // e.Ke, x.Cx, sse, bad flag, theta-gradient partials
    extern __shared__ double smem[];
    const int NT = a.NT, n = a.n, G = a.G, H = a.H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
    const int nwarps = blockDim.x >> 5;
    const int DW = nwarps / (G * H);                 // dimensions processed concurrently by the block
    const int g = warp % G, dslot = (warp / G) % DW, h = warp / (G * DW);
    const int TPS = (NT + H - 1) / H;
    const int T0 = h * TPS, T1 = (T0 + TPS < NT) ? T0 + TPS : NT;
    const size_t scr_doubles = (size_t)G * D * NT * 64;
    double* escr = a.scratch_in_smem ? smem : a.scratch + (size_t)blockIdx.x * 2 * scr_doubles;
    double* kscr = escr + scr_doubles;
    double* red = smem + (a.scratch_in_smem ? 2 * scr_doubles : 0);          // [G*8][D][H][RED]

    const long long chain = (long long)blockIdx.x * (G * 8) + g * 8 + gid;
    const bool cvalid = chain < a.n_chains;
    const double* xp = a.params + (cvalid ? chain : (long long)a.n_chains - 1) * a.pitch;
    double th[K];
#pragma unroll
    for (int i = 0; i < K; ++i) th[i] = xp[(size_t)n * D + i];
    const double inv_b1 = a.inv_beta[0], inv_b2 = a.inv_beta[1], inv_b3 = a.inv_beta[2];
    long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0, tk4 = 0, tk5 = 0;
#ifdef MAGI_DBG_FINE
    long long fine[4] = {0, 0, 0, 0};
#endif
    if (a.dbg) tk0 = clock64();

    // ---------------- P1: e = f - m~ x ----------------
    for (int d = dslot; d < D; d += DW) {
        double xw[WN];
#pragma unroll
        for (int i = 0; i < WN; ++i) xw[i] = 0.0;
        const double* xd = xp + (size_t)d * n;
        double* es = escr + ((size_t)(g * D + d) * NT) * 64 + lane;
        const double* ft0 = a.fragtab + ((size_t)(0 * D + d) * NT) * NCH * 32 + lane;
        const int s_begin = T0 - CH2, s_end = T1 - 1 + LAGT;
        double nx0, nx1;
        { const int t0 = 8 * s_begin + q, t1 = t0 + 4;
          nx0 = (t0 >= 0 && t0 < n) ? xd[t0] : 0.0; nx1 = (t1 >= 0 && t1 < n) ? xd[t1] : 0.0; }
        for (int s = s_begin; s <= s_end; ++s) {
#pragma unroll
            for (int i = 0; i < WN - 2; ++i) xw[i] = xw[i + 2];
            xw[WN - 2] = nx0;
            xw[WN - 1] = nx1;
            { const int t0 = 8 * (s + 1) + q, t1 = t0 + 4;           // next step's window feed
              nx0 = (t0 >= 0 && t0 < n) ? xd[t0] : 0.0; nx1 = (t1 >= 0 && t1 < n) ? xd[t1] : 0.0; }
            const int Ja = s - LAGT;
            if (Ja >= T0) {                                           // Ja < T1 by the loop bound
                const int t0 = 8 * Ja + q, t1 = t0 + 4;
                double xa[D], xb[D];
#pragma unroll
                for (int dd = 0; dd < D; ++dd) {
                    xa[dd] = (t0 < n) ? xp[(size_t)dd * n + t0] : 0.0;
                    xb[dd] = (t1 < n) ? xp[(size_t)dd * n + t1] : 0.0;
                }
                const double* f = ft0 + (size_t)Ja * NCH * 32;
                double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
#pragma unroll
                for (int hh = 0; hh < NCH; hh += 2) {
                    dmma884(m0, m1, xw[hh], __ldg(f + hh * 32));
                    dmma884(m2, m3, xw[hh + 1], __ldg(f + (hh + 1) * 32));
                }
                m0 += m2; m1 += m3;                                   // likelihoods.jl:129
                double e0 = 0.0, e1 = 0.0;
                if (t0 < n) e0 = M::f(d, xa, th) - m0;                // likelihoods.jl:130
                if (t1 < n) e1 = M::f(d, xb, th) - m1;
                es[(size_t)Ja * 64] = e0;
                es[(size_t)Ja * 64 + 32] = e1;
            }
        }
    }
    if (a.dbg) tk1 = clock64();
    __syncthreads();
    if (a.dbg) tk2 = clock64();

    // ---------------- P2: Ke = K~ e ----------------
    for (int d = dslot; d < D; d += DW) {
        const double* es = escr + ((size_t)(g * D + d) * NT) * 64 + lane;
        double* ks = kscr + ((size_t)(g * D + d) * NT) * 64 + lane;
        const double* ft2 = a.fragtab + ((size_t)(2 * D + d) * NT) * NCH * 32 + lane;
        double acc_eke = 0.0;
        for (int J = T0; J < T1; ++J) {
            const double* f = ft2 + (size_t)J * NCH * 32;
            double k0 = 0.0, k1 = 0.0, k2 = 0.0, k3 = 0.0;
#pragma unroll
            for (int hh = 0; hh < NCH; hh += 2) {
                const int c0 = 2 * J - HB + hh, c1 = c0 + 1;          // operand chunks (4 times each)
                const double o0 = (c0 >= 0 && c0 < 2 * NT) ? es[(size_t)(c0 >> 1) * 64 + (c0 & 1) * 32] : 0.0;
                const double o1 = (c1 >= 0 && c1 < 2 * NT) ? es[(size_t)(c1 >> 1) * 64 + (c1 & 1) * 32] : 0.0;
                dmma884(k0, k1, o0, __ldg(f + hh * 32));
                dmma884(k2, k3, o1, __ldg(f + (hh + 1) * 32));
            }
            k0 += k2; k1 += k3;                                       // likelihoods.jl:132
            ks[(size_t)J * 64] = k0;
            ks[(size_t)J * 64 + 32] = k1;
            acc_eke += es[(size_t)J * 64] * k0;                        // likelihoods.jl:146
            acc_eke += es[(size_t)J * 64 + 32] * k1;
        }
        acc_eke = quad_sum(acc_eke);
        if (q == 0) red[(((size_t)(g * 8 + gid) * D + d) * H + h) * RED + 0] = acc_eke;
    }
    if (a.dbg) tk3 = clock64();
    __syncthreads();
    if (a.dbg) tk4 = clock64();

    // ---------------- P3: Cx, m^T Ke, pointwise gradient ----------------
    for (int d = dslot; d < D; d += DW) {
        double xw[WN];
#pragma unroll
        for (int i = 0; i < WN; ++i) xw[i] = 0.0;
        double acc_xcx = 0.0, acc_sse = 0.0;
        double gth[K];
#pragma unroll
        for (int i = 0; i < K; ++i) gth[i] = 0.0;
        bool bad = false;
        double sigma_d;
        if (a.sigma_is_fixed) sigma_d = a.sigma_init[d];
        else {
            const double raw = xp[(size_t)n * D + K + d];
            const double ls = fmin(fmax(raw, -15.0), 15.0);           // interface.jl:200
            sigma_d = isnan(raw) ? raw : exp(ls);
        }
        const double inv_sig2 = 1.0 / (sigma_d * sigma_d);
        const double* xd = xp + (size_t)d * n;
        const double* yd = a.yobs + (size_t)d * n;
        const double* ks = kscr + ((size_t)(g * D + d) * NT) * 64 + lane;
        const double* ft1 = a.fragtab + ((size_t)(1 * D + d) * NT) * NCH * 32 + lane;
        const double* ft3 = a.fragtab + ((size_t)(3 * D + d) * NT) * NCH * 32 + lane;
        double* gout = (a.grad != nullptr && cvalid) ? a.grad + chain * a.pitch + (size_t)d * n : nullptr;
        const int s_begin = T0 - CH2, s_end = T1 - 1 + LAGT;
        double nx0, nx1;
        { const int t0 = 8 * s_begin + q, t1 = t0 + 4;
          nx0 = (t0 >= 0 && t0 < n) ? xd[t0] : 0.0; nx1 = (t1 >= 0 && t1 < n) ? xd[t1] : 0.0; }
        for (int s = s_begin; s <= s_end; ++s) {
#pragma unroll
            for (int i = 0; i < WN - 2; ++i) xw[i] = xw[i + 2];
            xw[WN - 2] = nx0;
            xw[WN - 1] = nx1;
            { const int t0 = 8 * (s + 1) + q, t1 = t0 + 4;
              nx0 = (t0 >= 0 && t0 < n) ? xd[t0] : 0.0; nx1 = (t1 >= 0 && t1 < n) ? xd[t1] : 0.0; }
            const int Jc = s - LAGT;
            if (Jc >= T0) {
#ifdef MAGI_DBG_FINE
                const long long f0 = clock64();
#endif
                const int t0 = 8 * Jc + q, t1 = t0 + 4;
                double xa[D], xb[D], wa[D], wb[D];
#pragma unroll
                for (int dd = 0; dd < D; ++dd) {
                    xa[dd] = (t0 < n) ? xp[(size_t)dd * n + t0] : 0.0;
                    xb[dd] = (t1 < n) ? xp[(size_t)dd * n + t1] : 0.0;
                    const double* wsrc = kscr + ((size_t)(g * D + dd) * NT + Jc) * 64 + lane;
                    wa[dd] = wsrc[0] * inv_b1;                         // likelihoods.jl:201
                    wb[dd] = wsrc[32] * inv_b1;
                }
                const double y0 = (t0 < n) ? __ldg(yd + t0) : 0.0, y1 = (t1 < n) ? __ldg(yd + t1) : 0.0;
                const double* f1 = ft1 + (size_t)Jc * NCH * 32;
                const double* f3 = ft3 + (size_t)Jc * NCH * 32;
                double c0 = 0.0, c1 = 0.0, u0 = 0.0, u1 = 0.0;
#ifdef MAGI_DBG_FINE
                const long long f1c = clock64();
#endif
#pragma unroll
                for (int hh = 0; hh < NCH; ++hh) {
                    const int ck = 2 * Jc - HB + hh;
                    const double ko = (ck >= 0 && ck < 2 * NT) ? ks[(size_t)(ck >> 1) * 64 + (ck & 1) * 32] : 0.0;
                    dmma884(c0, c1, xw[hh], __ldg(f1 + hh * 32));       // likelihoods.jl:133
                    dmma884(u0, u1, ko, __ldg(f3 + hh * 32));           // likelihoods.jl:192
                }
#ifdef MAGI_DBG_FINE
                const long long f2c = clock64() + (long long)(c0 * 0.0) + (long long)(u0 * 0.0);
#endif
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int t = u ? t1 : t0;
                    if (t < n) {
                        const double cx = u ? c1 : c0, mt = u ? u1 : u0;
                        const double* xv = u ? xb : xa;
                        const double* w = u ? wb : wa;
                        double xdv = 0.0, wd = 0.0;
#pragma unroll
                        for (int dd = 0; dd < D; ++dd) if (dd == d) { xdv = xv[dd]; wd = w[dd]; }
                        const double y = u ? y1 : y0;
                        const bool fin = isfinite(y);                  // likelihoods.jl:123
                        const double e0 = fin ? xdv - y : 0.0;
                        double gv = 0.0;
                        if (fin) gv -= (e0 * inv_sig2) * inv_b3;       // likelihoods.jl:179
                        gv -= cx * inv_b2;                             // likelihoods.jl:186
                        gv += mt * inv_b1;                             // likelihoods.jl:194
                        M::jx_col_sub(d, xv, th, w, gv);               // likelihoods.jl:214-216
                        M::jth_row_sub(d, xv, th, wd, gth);            // likelihoods.jl:219-221
                        acc_xcx += xdv * cx;                           // likelihoods.jl:150
                        acc_sse += e0 * e0;                            // likelihoods.jl:139,234
                        bad |= !isfinite(gv);
                        if (gout != nullptr) gout[t] = gv;
                    }
                }
#ifdef MAGI_DBG_FINE
                const long long f3c = clock64() + (long long)(acc_xcx * 0.0);
                fine[0] += f1c - f0; fine[1] += f2c - f1c; fine[2] += f3c - f2c; fine[3] += 1;
#endif
            }
        }
        acc_xcx = quad_sum(acc_xcx);
        acc_sse = quad_sum(acc_sse);
#pragma unroll
        for (int i = 0; i < K; ++i) gth[i] = quad_sum(gth[i]);
        const unsigned badm = __ballot_sync(0xffffffffu, bad);
        if (q == 0) {
            double* r = red + (((size_t)(g * 8 + gid) * D + d) * H + h) * RED;
            r[1] = acc_xcx;
            r[2] = acc_sse;
            r[3] = ((badm >> (gid * 4)) & 0xfu) ? 1.0 : 0.0;
#pragma unroll
            for (int i = 0; i < K; ++i) r[4 + i] = gth[i];
        }
    }
    if (a.dbg) tk5 = clock64();
    __syncthreads();
    if (a.dbg && lane == 0) {
        long long* o = a.dbg + ((size_t)blockIdx.x * nwarps + warp) * 8;
        o[0] = tk1 - tk0; o[1] = tk2 - tk1; o[2] = tk3 - tk2; o[3] = tk4 - tk3; o[4] = tk5 - tk4; o[5] = clock64() - tk5;
#ifdef MAGI_DBG_FINE
        o[5] = fine[0]; o[6] = fine[1]; o[7] = fine[2];
#endif
    }

    // ---------------- final: one thread per chain ----------------
    if (threadIdx.x < G * 8) {
        const long long c = (long long)blockIdx.x * (G * 8) + threadIdx.x;
        if (c < a.n_chains) {
            const double* cp = a.params + c * a.pitch;
            double* gp = a.grad ? a.grad + c * a.pitch : nullptr;
            const int nxt = n * D + K;
            const int P = a.P;
            if (a.sigma_invalid) {                                    // interface.jl:192-195
                a.ll[c] = -INFINITY;
                if (gp) for (int i = 0; i < P; ++i) gp[i] = NAN;
                return;
            }
            double ll = 0.0, prior = 0.0;
            double gsig[D], sig[D], gthf[K];
            bool bad = false;
#pragma unroll
            for (int i = 0; i < K; ++i) gthf[i] = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double eke = 0.0, xcx = 0.0, sse = 0.0;
                for (int hh = 0; hh < H; ++hh) {
                    const double* r = red + (((size_t)threadIdx.x * D + d) * H + hh) * RED;
                    eke += r[0]; xcx += r[1]; sse += r[2];
                    bad |= (r[3] != 0.0);
#pragma unroll
                    for (int i = 0; i < K; ++i) gthf[i] += r[4 + i];
                }
                double s;
                if (a.sigma_is_fixed) s = a.sigma_init[d];
                else {
                    const double raw = cp[nxt + d];
                    const double cl = fmin(fmax(raw, -15.0), 15.0);
                    s = isnan(raw) ? raw : exp(cl);
                    prior += isnan(raw) ? raw : cl;                   // interface.jl:206
                }
                sig[d] = s;
                const double s2 = s * s;
                const int nobs = a.nobs[d];
                double ll_obs = -0.5 * sse / s2;                      // likelihoods.jl:139
                if (nobs > 0) ll_obs -= 0.5 * nobs * log(2.0 * M_PI * s2);   // :141
                ll += ll_obs / a.beta[2];                             // :143
                ll += (-0.5 * eke) / a.beta[0];                       // :146-147
                ll += (-0.5 * xcx) / a.beta[1];                       // :150-151
                gsig[d] = (s > 0 && nobs > 0) ? (sse / s2 - nobs) / (s * a.beta[2]) : 0.0;   // :229-246
                bad |= !isfinite(gsig[d]);
            }
#pragma unroll
            for (int i = 0; i < K; ++i) bad |= !isfinite(gthf[i]);
            bad |= !isfinite(ll);
            if (bad) {                                                // interface.jl:222-226
                a.ll[c] = -INFINITY;
                if (gp) for (int i = 0; i < P; ++i) gp[i] = 0.0;
                return;
            }
            double total = ll;
            bool bad2 = false;
            double gls[D];
            if (!a.sigma_is_fixed) {
                total += prior;                                       // interface.jl:238
#pragma unroll
                for (int d = 0; d < D; ++d) { gls[d] = gsig[d] * sig[d] + 1.0; bad2 |= !isfinite(gls[d]); }   // :249-253
            }
            a.ll[c] = total;
            if (gp) {
                if (bad2) { for (int i = 0; i < P; ++i) gp[i] = 0.0; }   // interface.jl:260-264
                else {
#pragma unroll
                    for (int i = 0; i < K; ++i) gp[n * D + i] = gthf[i];
                    if (!a.sigma_is_fixed) {
#pragma unroll
                        for (int d = 0; d < D; ++d) gp[nxt + d] = gls[d];
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------------------------
bool model_dims(int model, int& D, int& K) {
    switch (model) {
    case MAGI_MODEL_FN: D = 2; K = 3; return true;
    case MAGI_MODEL_HES1: D = 3; K = 7; return true;
    case MAGI_MODEL_HES1LOG: D = 3; K = 7; return true;
    case MAGI_MODEL_HES1LOG_FIXG: D = 3; K = 6; return true;
    case MAGI_MODEL_HES1LOG_FIXF: D = 3; K = 6; return true;
    case MAGI_MODEL_HIV: D = 4; K = 9; return true;
    case MAGI_MODEL_PTRANS: D = 5; K = 6; return true;
    case MAGI_MODEL_LV: D = 2; K = 4; return true;
    default: return false;
    }
}

size_t banded_scratch_doubles_per_cta(int G, int D, int NT) { return (size_t)2 * G * D * NT * 64; }   // e and Ke

// Chooses chain-groups per block (G), time segments (H), concurrently processed dimensions (DW) and where the
// e / Ke scratch lives.  Preference: 8 warps per block and two blocks per SM (16 resident warps).
void banded_pick_config(int D, int K, int NT, int smem_limit, int& G, int& H, int& DW, int& scratch_in_smem, size_t& smem_bytes) {
    const int RED = 4 + K;
    DW = D < 8 ? D : 8;
    const int cand[6][2] = {{2, 2}, {2, 1}, {1, 2}, {1, 1}, {4, 1}, {4, 2}};
    int gmax = 4, hforce = 0;
    if (const char* e = getenv("MAGI_FORCE_G")) gmax = atoi(e);
    if (const char* e = getenv("MAGI_FORCE_H")) hforce = atoi(e);
    const bool force_global = getenv("MAGI_FORCE_GLOBAL_SCRATCH") != nullptr;
    for (int pass = 0; pass < 2 && !force_global; ++pass) {          // pass 0: two blocks per SM; pass 1: one block per SM
        for (int i = 0; i < 4; ++i) {
            int g = cand[i][0], hh = cand[i][1];
            if (g > gmax || (hforce && hh != hforce)) continue;
            if (hh > 1 && NT < 4 * hh) continue;
            if (g * DW * hh > 8) continue;
            size_t red = (size_t)g * 8 * D * hh * RED * sizeof(double);
            size_t scr = banded_scratch_doubles_per_cta(g, D, NT) * sizeof(double);
            size_t lim = pass == 0 ? (size_t)(smem_limit + 1024) / 2 - 1024 : (size_t)smem_limit;
            if (scr + red <= lim) { G = g; H = hh; scratch_in_smem = 1; smem_bytes = scr + red; return; }
        }
    }
    G = 2; H = (NT >= 8) ? 2 : 1;
    while (G * DW * H > 8 && H > 1) H >>= 1;
    while (G * DW * H > 8 && G > 1) G >>= 1;
    scratch_in_smem = 0;
    smem_bytes = (size_t)G * 8 * D * H * RED * sizeof(double);
}

template <int MODEL, int HB>
static cudaError_t launch_one(const BandedArgs& a, int DW, size_t smem_bytes, cudaStream_t st) {
    auto kern = banded_logpost_kernel<MODEL, HB>;
    static bool attr_set = false;    // per instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int threads = a.G * DW * a.H * 32;
    const int blocks = (a.n_chains + a.G * 8 - 1) / (a.G * 8);
    kern<<<blocks, threads, smem_bytes, st>>>(a);
    return cudaGetLastError();
}

template <int MODEL>
static cudaError_t launch_model(const BandedArgs& a, int HB, int DW, size_t smem_bytes, cudaStream_t st) {
    switch (HB) {
    case 0: return launch_one<MODEL, 0>(a, DW, smem_bytes, st);
    case 1: return launch_one<MODEL, 1>(a, DW, smem_bytes, st);
    case 2: return launch_one<MODEL, 2>(a, DW, smem_bytes, st);
    case 3: return launch_one<MODEL, 3>(a, DW, smem_bytes, st);
    case 4: return launch_one<MODEL, 4>(a, DW, smem_bytes, st);
    case 5: return launch_one<MODEL, 5>(a, DW, smem_bytes, st);
    case 6: return launch_one<MODEL, 6>(a, DW, smem_bytes, st);
    case 7: return launch_one<MODEL, 7>(a, DW, smem_bytes, st);
    case 8: return launch_one<MODEL, 8>(a, DW, smem_bytes, st);
    default: return cudaErrorInvalidValue;
    }
}

// `a.G`, `a.scratch_in_smem` must come from banded_pick_config; DW = blockDim warps / G.
cudaError_t launch_banded_cfg(int model, const BandedArgs& a, int HB, int DW, size_t smem_bytes, cudaStream_t st) {
    switch (model) {
    case MAGI_MODEL_FN: return launch_model<MAGI_MODEL_FN>(a, HB, DW, smem_bytes, st);
    case MAGI_MODEL_HES1: return launch_model<MAGI_MODEL_HES1>(a, HB, DW, smem_bytes, st);
    case MAGI_MODEL_LV: return launch_model<MAGI_MODEL_LV>(a, HB, DW, smem_bytes, st);
#ifndef MAGI_FAST_BUILD
    case MAGI_MODEL_HES1LOG: return launch_model<MAGI_MODEL_HES1LOG>(a, HB, DW, smem_bytes, st);
    case MAGI_MODEL_HES1LOG_FIXG: return launch_model<MAGI_MODEL_HES1LOG_FIXG>(a, HB, DW, smem_bytes, st);
    case MAGI_MODEL_HES1LOG_FIXF: return launch_model<MAGI_MODEL_HES1LOG_FIXF>(a, HB, DW, smem_bytes, st);
    case MAGI_MODEL_HIV: return launch_model<MAGI_MODEL_HIV>(a, HB, DW, smem_bytes, st);
    case MAGI_MODEL_PTRANS: return launch_model<MAGI_MODEL_PTRANS>(a, HB, DW, smem_bytes, st);
#endif
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace magi
