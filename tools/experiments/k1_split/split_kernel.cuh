// K1-split: the log posterior and its gradient (src/likelihoods.jl:43-257, src/logdensityproblems_interface.jl:176-267) for a
// batch of chains as THREE band-product launches with fused pointwise epilogues, chained by programmatic dependent launches,
// with the chain state and the intermediates in L2-resident planes stored in DMMA FRAGMENT ORDER (the chain-interleaved
// layout of SURVEY.md K7).  The two fused kernels (banded_kernel.cuh, flow_kernel.cuh) keep a chain group's whole evaluation
// on one SM and end at ~50 % DMMA utilisation inside their loops because every warp has as much hand-off, window and
// pointwise work as tensor work.  Here a warp does one thing:
//
//   it owns ONE (view, dimension, pair of output tiles) -- the B operand, 2 x NCH fragments, stays in its registers -- and
//   sweeps a range of chain groups: HB + 2 (+1) coalesced 16-byte loads of the A operand, 2 x NCH DMMAs, a short epilogue
//   on its 4 output times, 2 coalesced 16-byte stores.
//
// No queues, no barriers, no shared memory; 16 independent warps per SM hide each other's latencies, every SM of the machine
// has work whatever the batch size (units = views x D x pairs x chain-group ranges), and a stage boundary is a kernel boundary
// (the gradient's Jacobian terms need KE of all dimensions, K~ e needs e of the neighbouring tiles).
//
// Plane layout: F[d][chain group g][tile J][lane = 4 gid + q][slot s] = value(chain 8 g + gid, time 8 J + q + 4 s).  With the
// output slots of the band tables permuted the same way (the windowed kernel's fragment tables, banded_kernel.cu) the C
// fragment a warp stores IS the A fragment of chunks 2 J and 2 J + 1 of the next product: one 16-byte load per lane and
// tile, 512 contiguous bytes per warp (4 L1 wavefronts).  The first version of this route read the chain-contiguous state
// directly (lane (gid, q) -> 8 rows x 32 bytes per load instruction: 8 wavefronts for 256 bytes) and was bound by exactly
// that: 14 + 8 such loads per 24 DMMAs, 0.080 ms at 4096 chains whatever the band width (profiles/README.md).
// MT zero tiles on both sides of every row (written once, at allocation) stand in for the band's reach past the time axis.
//
//   prep   the chain state into fragment order; per chain: theta + its invariants (Ode::prepare: FP64 divisions), 1/sigma_d^2
//   S0     view m~ :  E  = f(x, theta) - m~ x              (likelihoods.jl:129-130)                               -> EF
//          view C~ :  G0 = -(x - y) / sigma^2 / beta3 - C~ x / beta2   (:179, :186); sums x.C~x, (x - y)^2       -> G0F, part
//   S1     view K~ :  KE = K~ e / beta1  (:132); sum e.KE                                                         -> KEF, part
//   S2     view m~^T: g = G0 + m~^T KE (:194) - J_x^T KE (:214-216); theta gradient (:219-221); non-finite flag   -> grad, part
//   fin    per chain: sums over the tile pairs in a fixed order, log density in the reference's order of accumulation, sigma
//          gradient, log-sigma transform, guards (interface.jl:192-264)                                            -> ll, grad
//
// 1/beta1 and 1/beta2 are folded into the K~ and C~ fragment tables (refresh_fragtab, magi_abi.cu).
#pragma once
#include <cmath>
#include <cstdlib>
#include "magi_internal.cuh"
#include "k1_primitives.cuh"
#include "dense_ode.cuh"

namespace magi {

struct SplitArgs {
    int n, D, K, P, n_chains, NP, NG, NGc, R, TR, MT, CW, sigma_is_fixed, sigma_invalid;
    long long pitch, plane;     // plane: doubles between two dimensions of a fragment-order plane (= NGc x TR x 64)
    const double* params; double* ll; double* grad;     // grad may be null (value only)
    const double* frag;         // permuted output slots, 1/beta folded: [4 views][D][NP][NCH][32 lanes][2 tiles]
    double *XF, *EF, *KEF, *G0F;
    double* part;               // [chain][D][NP][4 + K]: e.KE, x.C~x, sse, bad flag, theta-gradient partials
    double* cst;                // [group][CW = KX + D][8 chains]: theta and invariants, 1 / sigma_d^2
    const double* yobs; const int* nobs; const double* sigma_init;
    double beta3, inv_b3;
};

__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ double2 ldg_nc_f64x2(const double2* p) {
    double2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double sigma_of(const SplitArgs& a, const double* xp, int d) {
    if (a.sigma_is_fixed) return a.sigma_init[d];
    const double raw = xp[a.n * a.D + a.K + d];
    return isnan(raw) ? raw : exp(fmin(fmax(raw, -15.0), 15.0));      // interface.jl:200
}
// lane's double2 of tile J of (dimension d, chain group g) in a fragment-order plane
__device__ __forceinline__ size_t split_tile(const SplitArgs& a, int d, int g, int J, int lane) {
    return (size_t)d * a.plane + (((size_t)g * a.TR + (J + a.MT)) * 32 + lane) * 2;
}

// Blocks [0, tile_blocks): one warp per (chain group, tile) writes the tile of every dimension in fragment order (times past
// the end of the axis and chains past the end of the batch as zeros).  Remaining blocks: one thread per chain, constants.
template <int MODEL>
__global__ void __launch_bounds__(256) split_prep_kernel(const SplitArgs a, int tile_blocks) {
    constexpr int K = DenseOde<MODEL>::K, KX = DenseOde<MODEL>::KX;
    griddep_launch_dependents();
    const int n = a.n, D = a.D;
    if ((int)blockIdx.x < tile_blocks) {
        const int lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
        const long long u = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
        const int NTt = 2 * a.NP;
        if (u >= (long long)a.NG * NTt) return;
        const int g = (int)(u / NTt), J = (int)(u % NTt);
        const int c = 8 * g + gid, t0 = 8 * J + q, t1 = t0 + 4;
        const bool cok = c < a.n_chains;
        const double* xp = a.params + (long long)(cok ? c : 0) * a.pitch;
        for (int d = 0; d < D; ++d) {
            const double* xd = xp + (long long)d * n;
            double2 v;
            v.x = (cok && t0 < n) ? xd[t0] : 0.0;
            v.y = (cok && t1 < n) ? xd[t1] : 0.0;
            *reinterpret_cast<double2*>(a.XF + split_tile(a, d, g, J, lane)) = v;
        }
        return;
    }
    const int c = ((int)blockIdx.x - tile_blocks) * blockDim.x + threadIdx.x;
    if (c >= 8 * a.NG) return;
    double* cs = a.cst + ((size_t)(c >> 3) * a.CW) * 8 + (c & 7);
    if (c >= a.n_chains) {                       // ghost chains of the last group: finite constants
        for (int i = 0; i < a.CW; ++i) cs[(size_t)i * 8] = 0.0;
        return;
    }
    const double* xp = a.params + (long long)c * a.pitch;
    double th[KX];
#pragma unroll
    for (int i = 0; i < K; ++i) th[i] = xp[(size_t)n * D + i];
    DenseOde<MODEL>::prepare(th);
#pragma unroll
    for (int i = 0; i < KX; ++i) cs[(size_t)i * 8] = th[i];
    for (int d = 0; d < D; ++d) {
        const double s = sigma_of(a, xp, d);
        cs[(size_t)(KX + d) * 8] = 1.0 / (s * s);
    }
}

// Component i of a many-component model at this lane's time, from the five neighbours of dimension d kept in registers
// (Lorenz-96 couples x_{d-2} .. x_{d+2} only; DenseOde<L96> asks for nothing else).
struct NeighbourCache {
    double v[5];
    int d, D;
    __device__ __forceinline__ double operator()(int i) const {
        int r = i - d;
        if (r > 2) r -= D;
        if (r < -2) r += D;
        return r == -2 ? v[0] : (r == -1 ? v[1] : (r == 0 ? v[2] : (r == 1 ? v[3] : v[4])));
    }
};

// STAGE 0: views m~ and C~ on the chain state; 1: K~ on E; 2: m~^T on KE.  NW warps per block, one block per SM.
template <int MODEL, int HB, int STAGE, int NW>
__global__ void __launch_bounds__(NW * 32, 1) split_stage_kernel(const SplitArgs a) {
    using M = DenseOde<MODEL>;
    constexpr int NCH = 2 * HB + 2, K = M::K, KX = M::KX, NV = 4 + K, V = (STAGE == 0) ? 2 : 1;
    constexpr int OFF = HB & 1, JB = (HB + OFF) / 2, NTL = HB + 2 + OFF;     // first tile 2p - JB, NTL tiles cover the NCH + 2 chunks
    constexpr int SD = M::SD > 0 ? M::SD : 1;
    const int lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
    const int n = a.n, D = a.D, NP = a.NP;
    const long long w = (long long)blockIdx.x * NW + (threadIdx.x >> 5);
    const long long n_units = (long long)a.R * D * V * NP;
    if (w >= n_units) { griddep_launch_dependents(); return; }
    // consecutive warps: neighbouring pairs (and both views) of one (chain-group range, dimension): their A operands overlap in L1
    const int p = (int)(w % NP);
    const int vi = (int)((w / NP) % V);
    const int d = (int)((w / ((long long)NP * V)) % D);
    const int r = (int)(w / ((long long)NP * V * D));
    const int view = (STAGE == 0) ? vi : (STAGE == 1 ? 2 : 3);
    double2 fb[NCH];
    {
        const double2* fr = reinterpret_cast<const double2*>(a.frag) + (((size_t)view * D + d) * NP + p) * NCH * 32 + lane;
#pragma unroll
        for (int hh = 0; hh < NCH; ++hh) fb[hh] = ldg_nc_f64x2(fr + hh * 32);
    }
    griddep_launch_dependents();     // the next stage's blocks may take the SMs this grid leaves (they wait below for our completion)
    griddep_wait();                  // everything the previous launch wrote is visible from here on
    const int g0 = (int)((long long)r * a.NG / a.R), g1 = (int)((long long)(r + 1) * a.NG / a.R);
    const double* inplane = STAGE == 0 ? a.XF : (STAGE == 1 ? a.EF : a.KEF);
    auto ld_tile = [&](const double* plane, int dd, int g, int J) { return *reinterpret_cast<const double2*>(plane + split_tile(a, dd, g, J, lane)); };
    auto st_tile = [&](double* plane, int g, int J, double v0, double v1) { *reinterpret_cast<double2*>(plane + split_tile(a, d, g, J, lane)) = make_double2(v0, v1); };

    // Software pipeline: the A tiles of group g + 1 and the epilogue inputs of group g are requested BEFORE the DMMAs of group g,
    // so a warp's memory latency overlaps its own tensor work (with 2-4 warps per SM sub-partition nothing else would hide it).
    auto ld_a = [&](int g, double2* tl) {
        const double2* src = reinterpret_cast<const double2*>(inplane + split_tile(a, d, g, 2 * p - JB, lane));
#pragma unroll
        for (int i = 0; i < NTL; ++i) tl[i] = src[i * 32];
    };
    constexpr int NX = (M::SD > 0) ? SD : 5;          // state components an epilogue needs at a time point
    auto xdim = [&](int i) { return (M::SD > 0) ? i : (d + i - 2 + D) % D; };
    double2 tl[NTL];
    if (g0 < g1) ld_a(g0, tl);

#pragma unroll 1
    for (int g = g0; g < g1; ++g) {
        const int c = 8 * g + gid;
        const bool cok = c < a.n_chains;
        double av[NCH + 2];
#pragma unroll
        for (int k = 0; k < NCH + 2; ++k) av[k] = ((k + OFF) & 1) ? tl[(k + OFF) >> 1].y : tl[(k + OFF) >> 1].x;
        if (g + 1 < g1) ld_a(g + 1, tl);
        // epilogue inputs of this group
        const double* cs = a.cst + ((size_t)g * a.CW) * 8 + gid;
        double th[KX];
        double2 xo[2][NX], kev[2][NX], g0v[2];
        double yv[2][2], inv_sig2 = 0.0;
        if constexpr (STAGE == 0 || STAGE == 2) {
            if (STAGE == 2 || vi == 0) {
#pragma unroll
                for (int i = 0; i < KX; ++i) th[i] = cs[i * 8];
#pragma unroll
                for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                    for (int i = 0; i < NX; ++i) xo[tt][i] = ld_tile(a.XF, xdim(i), g, 2 * p + tt);
            }
        }
        if constexpr (STAGE == 2) {
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                g0v[tt] = ld_tile(a.G0F, d, g, 2 * p + tt);
#pragma unroll
                for (int i = 0; i < NX; ++i) kev[tt][i] = ld_tile(a.KEF, xdim(i), g, 2 * p + tt);
            }
        }
        if constexpr (STAGE == 0) {
            if (vi == 1) {
                inv_sig2 = cs[(KX + d) * 8];
#pragma unroll
                for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const int t = 8 * (2 * p + tt) + q + 4 * s;
                        yv[tt][s] = a.yobs[(size_t)d * n + (t < n ? t : 0)];
                    }
            }
        }
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int hh = 0; hh < NCH; ++hh) {
            dmma884(acc[0][0], acc[0][1], av[hh], fb[hh].x);
            dmma884(acc[1][0], acc[1][1], av[hh + 2], fb[hh].y);
        }
        // this lane's outputs: acc[tt][s] at time 8 (2p + tt) + q + 4 s; the operand at the same place is av[HB + 2 tt + s]
        double* prt = a.part + (((size_t)(cok ? c : 0) * D + d) * NP + p) * NV;
        // component i of the state / of KE at (tt, s): registers for small models, the five coupled neighbours for Lorenz-96
        auto state_at = [&](const double2 (&src)[2][NX], int tt, int s, double* out) {
#pragma unroll
            for (int i = 0; i < NX; ++i) out[i] = s ? src[tt][i].y : src[tt][i].x;
        };

        if constexpr (STAGE == 0) {
            if (vi == 0) {                                   // E = f(x, theta) - m~ x
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) {
                    double e[2];
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        if constexpr (M::SD > 0) {
                            double xa[NX];
                            state_at(xo, tt, s, xa);
                            e[s] = M::f(d, [&](int i) { return xa[i]; }, th, D) - acc[tt][s];        // likelihoods.jl:130
                        } else {
                            NeighbourCache x;
                            x.d = d; x.D = D;
                            state_at(xo, tt, s, x.v);
                            e[s] = M::f(d, x, th, D) - acc[tt][s];
                        }
                        if (8 * (2 * p + tt) + q + 4 * s >= n) e[s] = 0.0;
                    }
                    st_tile(a.EF, g, 2 * p + tt, e[0], e[1]);
                }
            } else {                                         // G0 = observation term - C~ x / beta2; sums x.C~x and (x - y)^2
                double xcx = 0.0, sse = 0.0;
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) {
                    double gq[2];
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const int t = 8 * (2 * p + tt) + q + 4 * s;
                        const double xv = av[HB + 2 * tt + s], y = yv[tt][s], cx = acc[tt][s];
                        const bool fin = t < n && isfinite(y);
                        const double e0 = fin ? xv - y : 0.0;
                        double gv = 0.0;
                        if (fin) gv -= (e0 * inv_sig2) * a.inv_b3;                // likelihoods.jl:179
                        gv -= cx;                                                 // :186 (1/beta2 in the table)
                        xcx += xv * cx;
                        sse += e0 * e0;
                        gq[s] = gv;
                    }
                    st_tile(a.G0F, g, 2 * p + tt, gq[0], gq[1]);
                }
                xcx = quad_sum(xcx);
                sse = quad_sum(sse);
                if (cok && q == 0) { prt[1] = xcx; prt[2] = sse; }
            }
        } else if constexpr (STAGE == 1) {                   // KE = K~ e / beta1; sum e.KE
            double eke = 0.0;
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                eke += av[HB + 2 * tt] * acc[tt][0];
                eke += av[HB + 2 * tt + 1] * acc[tt][1];
                st_tile(a.KEF, g, 2 * p + tt, acc[tt][0], acc[tt][1]);
            }
            eke = quad_sum(eke);
            if (cok && q == 0) prt[0] = eke;
        } else {                                             // gradient with respect to the states, theta-gradient partials
            double gth[K];
#pragma unroll
            for (int i = 0; i < K; ++i) gth[i] = 0.0;
            double* gp = (a.grad && cok) ? a.grad + (long long)c * a.pitch + (long long)d * n : nullptr;
            bool bad = false;
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int t = 8 * (2 * p + tt) + q + 4 * s;
                    if (t < n) {
                        double gv = (s ? g0v[tt].y : g0v[tt].x) + acc[tt][s];        // likelihoods.jl:194
                        const double ked = av[HB + 2 * tt + s];                      // KE of this dimension: the operand itself
                        if constexpr (M::SD > 0) {
                            double xa[NX], wa[NX];
                            state_at(xo, tt, s, xa);
                            state_at(kev, tt, s, wa);
                            auto x = [&](int i) { return xa[i]; };
                            M::jx_col_sub(d, x, [&](int i) { return wa[i]; }, th, D, gv);      // :214-216
                            M::jth_row_sub(d, x, th, D, ked, gth);                             // :219-221
                        } else {
                            NeighbourCache x, wv;
                            x.d = wv.d = d; x.D = wv.D = D;
                            state_at(xo, tt, s, x.v);
                            state_at(kev, tt, s, wv.v);
                            M::jx_col_sub(d, x, wv, th, D, gv);
                            M::jth_row_sub(d, x, th, D, ked, gth);
                        }
                        bad |= a.grad && !isfinite(gv);       // value-only calls judge the log density alone (interface.jl:155-160)
                        if (gp) gp[t] = gv;
                    }
                }
            }
            double fl = quad_sum(bad ? 1.0 : 0.0);
#pragma unroll
            for (int i = 0; i < K; ++i) gth[i] = quad_sum(gth[i]);
            if (cok && q == 0) {
                prt[3] = fl;
#pragma unroll
                for (int i = 0; i < K; ++i) prt[4 + i] = gth[i];
            }
        }
    }
}

// One warp per chain: the pair partials summed in a fixed order, then the reference's assembly (interface.jl:192-264).
template <int MODEL>
__global__ void __launch_bounds__(128) split_finalize_kernel(const SplitArgs a) {
    constexpr int K = DenseOde<MODEL>::K, NV = 4 + K;
    extern __shared__ double sm[];
    griddep_launch_dependents();
    griddep_wait();
    const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 4 + wp, n = a.n, D = a.D, NP = a.NP, P = a.P;
    if (c >= a.n_chains) return;
    double* red = sm + (size_t)wp * (D * NV + 4 * D);
    double* fin = red + D * NV;       // [D][4]: ll_obs / beta3, -e.KE / 2, -x.C~x / 2, d / d log sigma
    const double* xp = a.params + (long long)c * a.pitch;
    double* gp = a.grad ? a.grad + (long long)c * a.pitch : nullptr;
    const int nxt = n * D + K;
    if (a.sigma_invalid) {                                            // interface.jl:192-195
        if (lane == 0) a.ll[c] = -INFINITY;
        if (gp) for (int i = lane; i < P; i += 32) gp[i] = NAN;
        return;
    }
    for (int i = lane; i < D * NV; i += 32) {
        const int d = i / NV, j = i - d * NV;
        const double* src = a.part + (((size_t)c * D + d) * NP) * NV + j;
        double s = 0.0;
        for (int p = 0; p < NP; ++p) s += src[(size_t)p * NV];
        red[i] = s;
    }
    __syncwarp();
    int flag = 0;
    for (int d = lane; d < D; d += 32) {
        const double* r = red + d * NV;
        const double s = sigma_of(a, xp, d);
        const double s2 = s * s, sse = r[2];
        const int nobs = a.nobs[d];
        double ll_obs = -0.5 * sse / s2;                              // likelihoods.jl:139
        if (nobs > 0) ll_obs -= 0.5 * nobs * log(2.0 * M_PI * s2);   // :141
        fin[4 * d + 0] = ll_obs / a.beta3;                            // :143
        fin[4 * d + 1] = -0.5 * r[0];                                 // :146-147 (1/beta1 in the K~ table)
        fin[4 * d + 2] = -0.5 * r[1];                                 // :150-151 (1/beta2 in the C~ table)
        const double gsig = (s > 0 && nobs > 0) ? (sse / s2 - nobs) / (s * a.beta3) : 0.0;   // :229-246
        const double gls = gsig * s + 1.0;                            // interface.jl:249-253
        fin[4 * d + 3] = gls;
        if (r[3] != 0.0 || (gp && !isfinite(gsig))) flag |= 1;
        if (!a.sigma_is_fixed && !isfinite(gls)) flag |= 2;
    }
    __syncwarp();
    double ll = 0.0, gth[K];
    if (lane == 0) {
        double prior = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) gth[i] = 0.0;
        for (int d = 0; d < D; ++d) {                                 // the reference's order of accumulation
            ll += fin[4 * d + 0]; ll += fin[4 * d + 1]; ll += fin[4 * d + 2];
#pragma unroll
            for (int i = 0; i < K; ++i) gth[i] += red[d * NV + 4 + i];
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
        if (bad) flag |= 1;
        if (!a.sigma_is_fixed) ll += prior;
    }
    flag = __reduce_or_sync(0xffffffffu, flag);
    if (flag & 1) {                                                    // interface.jl:222-226
        if (lane == 0) a.ll[c] = -INFINITY;
        if (gp) for (int i = lane; i < P; i += 32) gp[i] = 0.0;
        return;
    }
    if (lane == 0) a.ll[c] = ll;
    if (gp) {
        if (flag & 2) { for (int i = lane; i < P; i += 32) gp[i] = 0.0; }   // interface.jl:260-264
        else {
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < K; ++i) gp[n * D + i] = gth[i];
            }
            if (!a.sigma_is_fixed) for (int d = lane; d < D; d += 32) gp[nxt + d] = fin[4 * d + 3];
        }
    }
}

template <class Kern, class... Extra>
static cudaError_t launch_pdl(Kern kern, int blocks, int threads, size_t smem, cudaStream_t st, bool dependent, const SplitArgs& a, Extra... extra) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("MAGI_SPLIT_NOPDL") != nullptr;      // development knob
    cfg.attrs = attr; cfg.numAttrs = (dependent && !no_pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, a, extra...);
}

#ifndef MAGI_SPLIT_NW
#define MAGI_SPLIT_NW 8
#endif
constexpr int kSplitWarps = MAGI_SPLIT_NW;

template <int MODEL, int HB>
static cudaError_t split_launch_hb(SplitArgs a, int sm_count, cudaStream_t st, long long* launches) {
    constexpr int NW = kSplitWarps;
    const int NV = 4 + DenseOde<MODEL>::K;
    cudaError_t e;
    // prep is NOT a dependent launch: it overwrites cst / the planes the previous evaluation's kernels may still be reading
    const int tile_blocks = (int)(((long long)a.NG * 2 * a.NP + 7) / 8), cst_blocks = (8 * a.NG + 255) / 256;
    e = launch_pdl(split_prep_kernel<MODEL>, tile_blocks + cst_blocks, 256, 0, st, false, a, tile_blocks);
    if (e != cudaSuccess) return e;
    auto stage = [&](auto kern, int V) {
        const long long U = (long long)a.D * V * a.NP;
        long long R = ((long long)sm_count * NW) / U;
        if (R < 1) R = 1;
        if (R > a.NG) R = a.NG;
        a.R = (int)R;
        const long long blocks = (R * U + NW - 1) / NW;
        return launch_pdl(kern, (int)blocks, NW * 32, 0, st, true, a);
    };
    e = stage(split_stage_kernel<MODEL, HB, 0, NW>, 2);
    if (e != cudaSuccess) return e;
    e = stage(split_stage_kernel<MODEL, HB, 1, NW>, 1);
    if (e != cudaSuccess) return e;
    e = stage(split_stage_kernel<MODEL, HB, 2, NW>, 1);
    if (e != cudaSuccess) return e;
    const size_t fin_smem = sizeof(double) * 4 * ((size_t)a.D * NV + 4 * a.D);
    if (fin_smem > 48 * 1024) {
        static PerDeviceOnce once;
        if (once.need()) {
            e = cudaFuncSetAttribute(split_finalize_kernel<MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem);
            if (e != cudaSuccess) return e;
        }
    }
    e = launch_pdl(split_finalize_kernel<MODEL>, (a.n_chains + 3) / 4, 128, fin_smem, st, true, a);
    if (e != cudaSuccess) return e;
    *launches += 5;
    return cudaSuccess;
}

template <int MODEL>
cudaError_t split_launch_model(const SplitArgs& a, int HB, int sm_count, cudaStream_t st, long long* launches) {
    switch (HB) {
    case 0: return split_launch_hb<MODEL, 0>(a, sm_count, st, launches);
    case 1: return split_launch_hb<MODEL, 1>(a, sm_count, st, launches);
    case 2: return split_launch_hb<MODEL, 2>(a, sm_count, st, launches);
    case 3: return split_launch_hb<MODEL, 3>(a, sm_count, st, launches);
    case 4: return split_launch_hb<MODEL, 4>(a, sm_count, st, launches);
    case 5: return split_launch_hb<MODEL, 5>(a, sm_count, st, launches);
    case 6: return split_launch_hb<MODEL, 6>(a, sm_count, st, launches);
    case 7: return split_launch_hb<MODEL, 7>(a, sm_count, st, launches);
    case 8: return split_launch_hb<MODEL, 8>(a, sm_count, st, launches);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace magi
