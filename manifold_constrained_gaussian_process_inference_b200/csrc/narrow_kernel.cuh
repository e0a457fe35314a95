// K1-narrow: the log posterior and its gradient (src/likelihoods.jl:43-257, src/logdensityproblems_interface.jl:176-267) for
// band half-widths b <= 4 -- the regime where an evaluation is HBM / FP64 balanced (FN n=201, b=2 is 28 kFLOP
// against 6.5 kB) and the DMMA tiling wastes 44-80 % of its multiplications on the parallelogram corners (a tile pair costs
// 16 columns for 2b+1 useful ones) while its per-step fixed cost does not shrink with the band.  Plain FP64 FMAs instead:
//
//   * ONE THREAD PER CHAIN sweeps the time axis once.  The three operand windows (x: 4b+1 values per component, e and Ke:
//     2b+1 each) live in registers with static indices: the sweep is a software pipeline of three stages, at step s
//         stage 1 (t1 = s - b):   mx = m~ x,  e = f(x, theta) - mx                       (likelihoods.jl:129-130)
//         stage 2 (t2 = s - 2b):  Ke = K~ e / beta1, sum e.Ke                            (:132, :146)
//         stage 3 (t3 = s - 3b):  Cx = C~ x / beta2, mt = m~^T Ke, gradient, sums        (:133, :150, :179-221)
//     TB steps are unrolled so that the windows move TB places at a time (one register move per value and TB steps);
//   * the band coefficients of a step are the same for every chain: a per-step table [s][m~@t1 | C~@t3 | K~@t2 | m~^T@t3 | y@t3]
//     (built once per handle, zero where a row or a column falls off the time axis, 1/beta folded in) is staged per chunk of
//     16 steps in shared memory and read with warp-uniform (broadcast) 16-byte loads (two wavefronts per coefficient pair);
//   * the chain state is chain-contiguous in HBM (the host layout): a chunk of 16 times x up to 256 chains comes in as 128-byte
//     row segments by cp.async (zero fill past the end of the axis / batch), double-buffered against the compute of the previous
//     chunk, and is transposed through shared memory (odd row pitch: conflict-free both ways); the gradient goes back the same
//     way, so every state byte crosses HBM exactly once in each direction.
// Roofline: HBM for b <= 2, FP64 (DFMA, the same pipe as DMMA on this chip) for b = 3, 4.  Measured (FN n=201, 65 536 chains, one
// B200, bench.py -> smallband): b = 1 / 2 / 4 0.185 / 0.214 / 0.281 ms against 0.414 ms for the windowed DMMA kernel = 0.36 / 0.31 / 0.24
// of the HBM peak.  What holds it there (profiles/narrow_b2_r02_ncu_summary.txt; timing-only ablations): 7-8 warps per SM (the
// registers of the windows), issue slots 39 % busy, FP64 pipe 33 %; the copies in and out cost ~0.08 ms of a pass and do not
// overlap with the compute when every warp issues its share of them between its steps; the coefficient loads cost nothing
// measurable.  Hence the second, warp-specialised organisation of the copies (template flag WS below).  Tried and slower or
// without effect: DESIGN.md section 4, K1-narrow.
#pragma once
#include <cmath>
#include "magi_internal.cuh"
#include "ode_models.cuh"

namespace magi {

constexpr int kNarrowMaxChains = 256;   // chains (= threads) per block: a multiple of 32 chosen per call (whole waves of one block per SM)
constexpr int kNarrowTCH = 16;          // steps the per-step table is padded to (the largest staged chunk)

struct NarrowArgs {
    int n, P, n_chains, sigma_is_fixed, sigma_invalid, CS, n_steps_pad, n_tiles;
    long long pitch;
    const double* params; double* ll; double* grad;     // grad may be null (value only)
    const double* steptab;      // [n_steps_pad][CS]: per step m~@t1 [D][W], C~@t3 [D][W], K~@t2 [D][W], m~^T@t3 [D][WP], (y, weight)@t3 [D][2]
    const int* nobs; const double* sigma_init;
    double beta3, inv_b3;
};

// per step: 4 D coefficient rows of WP = 2b + 2 doubles (2b + 1 used: even length, every row 16-byte aligned), then (y, weight)[D]
inline int narrow_cs(int D, int b) { return 4 * D * (2 * b + 2) + 2 * D; }
inline int narrow_steps_pad(int n, int b) { return (n + 3 * b + kNarrowTCH - 1) / kNarrowTCH * kNarrowTCH; }

// band tables are diagonal-major: T[(b + j - i) n + i] = A[i, j]
__global__ void build_steptab_kernel(const double* __restrict__ band_cinv, const double* __restrict__ band_mphi, const double* __restrict__ band_kinv,
                                     const double* __restrict__ yobs, double* __restrict__ steptab, int n, int b, int D, int CS, int n_steps_pad,
                                     double scale_c, double scale_k) {
    const int W = 2 * b + 1, WP = W + 1;
    const size_t total = (size_t)n_steps_pad * CS, tab = (size_t)W * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(idx / CS), r = (int)(idx % CS);
        double v = 0.0;
        if (r < 4 * D * WP && r % WP < W) {
            const int view = r / (D * WP), d = (r / WP) % D, o = r % WP - b;
            const int t = s - (view == 0 ? b : (view == 2 ? 2 * b : 3 * b));     // output time of this view at step s
            const int j = t + o;
            if (t >= 0 && t < n && j >= 0 && j < n) {
                if (view == 0) v = band_mphi[d * tab + (size_t)(b + o) * n + t];
                else if (view == 1) v = band_cinv[d * tab + (size_t)(b + o) * n + t] * scale_c;
                else if (view == 2) v = band_kinv[d * tab + (size_t)(b + o) * n + t] * scale_k;
                else v = band_mphi[d * tab + (size_t)(b - o) * n + j];           // m~[j, t]
            }
        } else if (r >= 4 * D * WP) {                         // (y, 1) where dimension d is observed at t3, (0, 0) elsewhere
            const int d = (r - 4 * D * WP) >> 1, t = s - 3 * b;
            const double y = (t >= 0 && t < n) ? yobs[(size_t)d * n + t] : NAN;
            v = isfinite(y) ? (((r - 4 * D * WP) & 1) ? 1.0 : y) : 0.0;
        }
        steptab[idx] = v;
    }
}

__device__ __forceinline__ void cp_async8_zfill(double* smem_dst, const double* gsrc, bool pred) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 8 : 0;                                   // src-size 0: the destination is zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// TCH: steps per staged chunk.  Two organisations of the copies around the same arithmetic (a chain's result has the same bits in
// both, so the choice between them is a matter of speed only):
//   WS = false  every thread is a chain and issues its share of the copies between its steps: up to 256 chains per block, one block
//               per SM, chunks of 16 steps, one gradient tile -- the fastest for batches of many waves at b <= 2;
//   WS = true   WARP-SPECIALISED: the last warp of the block only moves data -- the next chunk's state and coefficient rows in, the
//               previous chunk's gradient tile out -- while up to three compute warps (96 chains) work on the current chunk; one block
//               barrier per chunk of 8 steps, two gradient tiles, two blocks per SM.  One staging warp keeps up with three compute
//               warps, not with seven (7 + 1 warps: 0.30 ms at every b).  Faster for mid-size batches and for b = 3, 4.
// Timing-only ablations of WS = false (FN n=201, 65 536 chains, ms with / without the copies): b = 2 0.210 / 0.128, b = 0 0.158 / 0.080;
// without the coefficient loads 0.206 -- the copies, issued by the computing warps, are what does not overlap.
template <int MODEL, int B, int TB, int TCH, bool WS>
__global__ void __launch_bounds__(WS ? 128 : kNarrowMaxChains, WS ? 2 : 1) narrow_logpost_kernel(const NarrowArgs a) {
    using M = Ode<MODEL>;
    constexpr int D = M::D, K = M::K, KX = M::KX, W = 2 * B + 1, WP = W + 1;
    constexpr int XW = 4 * B + TB, EW = 2 * B + TB, CS = 4 * D * WP + 2 * D;        // = narrow_cs(D, B)
    constexpr int NGS = WS ? 2 : 1;                                                 // gradient tiles
    static_assert(TCH % TB == 0, "chunk / unroll");
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, NC = blockDim.x - (WS ? 32 : 0), CP = NC + 1, n = a.n;     // NC chains per block; CP: odd row pitch of the tiles
    double* xs0 = sm;                                  // [2 buffers][D][TCH][CP] state of a chunk's times
    double* gs0 = xs0 + 2 * D * TCH * CP;              // [NGS buffers][D][TCH][CP] gradient of a chunk's stage-3 times
    double* cf0 = gs0 + NGS * D * TCH * CP;            // [2 buffers][TCH][CS] ((2 + NGS) D TCH CP doubles before it: even, 16-byte aligned)
    const long long c0 = (long long)blockIdx.x * NC;
    const int n_chunks = a.n_steps_pad / TCH;
    const bool want_grad = a.grad != nullptr;
    // The copies are done by SN "stagers": every thread (WS = false) or the lanes of the staging warp.  Element idx = sid + k SN of
    // a tile: time tt = idx % TCH and dimension (idx / TCH) % D do not depend on k (SN is a multiple of TCH D), the chain advances
    // by SN / (TCH D) per k: one running pointer per stager, no index arithmetic in the loops.
    static_assert(TCH * D <= 32 && 32 % (TCH * D) == 0, "tile mapping");
    const int SN = WS ? 32 : NC, sid = WS ? tid - NC : tid;        // (sid < 0: a compute thread of the warp-specialised kernel, copies nothing)
    const int sidc = sid < 0 ? 0 : sid;
    const int m_tt = sidc % TCH, m_d = (sidc / TCH) % D, m_cl = sidc / (TCH * D);
    const int m_step = SN / (TCH * D), m_iter = WS ? NC * D * TCH / 32 : TCH * D;      // (copies per stager and tile: a constant when every thread copies)
    const int m_soff = (m_d * TCH + m_tt) * CP + m_cl;
    // asynchronous copy of chunk ch (state of times [s0, s0 + TCH) as row segments of TCH doubles, zero past the end of the axis / of
    // the batch; the chunk's coefficient rows) into buffer ch & 1
    auto stage_in = [&](int ch) {
        const int s0 = ch * TCH, t = s0 + m_tt;
        double* xs = xs0 + (ch & 1) * D * TCH * CP + m_soff;
        const bool tok = t < n;
        const long long rows_left = (long long)a.n_chains - c0 - m_cl;       // chains from this stager's first one to the end of the batch
        const long long sstep = (long long)m_step * a.pitch;
        const double* src = a.params + (c0 + m_cl) * a.pitch + (long long)m_d * n + (tok ? t : 0);
#pragma unroll 4
        for (int k = 0; k < m_iter; ++k) {
            const bool ok = tok && (long long)k * m_step < rows_left;
            cp_async8_zfill(xs + k * m_step, ok ? src : a.params, ok);
            src += sstep;
        }
        const double* csrc = a.steptab + (size_t)s0 * CS;
        double* dst = cf0 + (ch & 1) * TCH * CS;
        for (int idx = sid; idx < TCH * CS / 2; idx += SN) cp_async16(dst + 2 * idx, csrc + 2 * idx);
        cp_async_commit();
    };
    // gradient of chunk ch (times [s0 - 3B, s0 + TCH - 3B)) from its tile to global memory
    auto stage_out = [&](int ch) {
        const int t = ch * TCH + m_tt - 3 * B;
        if (t < 0 || t >= n) return;
        const long long rows_left = (long long)a.n_chains - c0 - m_cl;
        const long long sstep = (long long)m_step * a.pitch;
        double* dst = a.grad + (c0 + m_cl) * a.pitch + (long long)m_d * n + t;
        const double* g = gs0 + (ch & (NGS - 1)) * D * TCH * CP + m_soff;
        if constexpr (WS) {
#pragma unroll 1
            for (int k0 = 0; k0 < m_iter; k0 += 8) {
                double v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = (k0 + k < m_iter) ? g[(k0 + k) * m_step] : 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (k0 + k < m_iter && (long long)(k0 + k) * m_step < rows_left) dst[(k0 + k) * sstep] = v[k];
            }
        } else {
#pragma unroll 4
            for (int k = 0; k < m_iter; ++k) {
                if ((long long)k * m_step < rows_left) *dst = g[k * m_step];
                dst += sstep;
            }
        }
    };
    if constexpr (WS) {
        if (tid >= NC) {
            // ---------------- the staging warp ----------------
            stage_in(0);
            cp_async_wait_all();
            __syncthreads();
#pragma unroll 1
            for (int ch = 0; ch < n_chunks; ++ch) {
                if (ch + 1 < n_chunks) stage_in(ch + 1);       // its buffer was last read in iteration ch - 1
                if (ch >= 1 && want_grad) stage_out(ch - 1);    // written in iteration ch - 1, rewritten in iteration ch + 1
                cp_async_wait_all();
                __syncthreads();
            }
            if (want_grad) stage_out(n_chunks - 1);
            __syncthreads();       // every x-gradient store of the block is issued before a guard of the compute warps overwrites a chain's row
            return;
        }
    } else {
        stage_in(0);
    }

    // ---------------- one thread per chain ----------------
    const long long c = c0 + tid;
    const bool valid = c < a.n_chains;
    const double* xp = a.params + (valid ? c : (long long)a.n_chains - 1) * a.pitch;
    const int nxt = n * D + K;

    double th[KX];
#pragma unroll
    for (int i = 0; i < K; ++i) th[i] = xp[(size_t)n * D + i];
    M::prepare(th);
    double sig[D], obs_scale[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        double s;
        if (a.sigma_is_fixed) s = a.sigma_init[d];
        else {
            const double raw = xp[nxt + d];
            s = isnan(raw) ? raw : exp(fmin(fmax(raw, -15.0), 15.0));     // interface.jl:200
        }
        sig[d] = s;
        obs_scale[d] = (1.0 / (s * s)) * a.inv_b3;
    }
    double xw[D][XW], ew[D][EW], kw[D][EW];
#pragma unroll
    for (int d = 0; d < D; ++d) {
#pragma unroll
        for (int i = 0; i < XW; ++i) xw[d][i] = 0.0;
#pragma unroll
        for (int i = 0; i < EW; ++i) { ew[d][i] = 0.0; kw[d][i] = 0.0; }
    }
    double eke[D], xcx[D], sse[D], gth[K];
#pragma unroll
    for (int d = 0; d < D; ++d) { eke[d] = 0.0; xcx[d] = 0.0; sse[d] = 0.0; }
#pragma unroll
    for (int i = 0; i < K; ++i) gth[i] = 0.0;
    bool bad = false;

    if constexpr (WS) __syncthreads();       // chunk 0 is in shared memory
#pragma unroll 1
    for (int ch = 0; ch < n_chunks; ++ch) {
        const int s0 = ch * TCH;
        if constexpr (!WS) {
            cp_async_wait_all();
            __syncthreads();                 // chunk ch is in shared memory; everybody is done with chunk ch - 1 (its buffer, its gradient tile)
            if (ch + 1 < n_chunks) stage_in(ch + 1);
        }
        const double* xs = xs0 + (ch & 1) * D * TCH * CP;
        const double* cf = cf0 + (ch & 1) * TCH * CS;
        double* gs = gs0 + (ch & (NGS - 1)) * D * TCH * CP;
#pragma unroll 1
        for (int sb = 0; sb < TCH / TB; ++sb) {
            // Stage-major over the TB unrolled steps: the TB band products of a stage are independent dependency chains (a
            // step-major body is one long chain per step, and its shared-memory stores keep the next step's loads behind them).
            const double* cfb = cf + sb * TB * CS;
            // coefficient o of row (view v, dimension d) of step u: rows are 16-byte aligned, one LDS.128 per two coefficients
            auto coef = [&](const double2* row, int o) { const double2 p = row[o >> 1]; return (o & 1) ? p.y : p.x; };
#pragma unroll
            for (int u = 0; u < TB; ++u)
#pragma unroll
                for (int d = 0; d < D; ++d) xw[d][4 * B + u] = xs[(d * TCH + sb * TB + u) * CP + tid];
            // ---- stage 1: t1 = s - B ----
#pragma unroll
            for (int u = 0; u < TB; ++u) {
                const double* cfs = cfb + u * CS;
                const int t1 = s0 + sb * TB + u - B;
                double xa[D];
#pragma unroll
                for (int d = 0; d < D; ++d) xa[d] = xw[d][3 * B + u];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const double2* r0 = reinterpret_cast<const double2*>(cfs + (0 * D + d) * WP);
                    double mx = coef(r0, 0) * xw[d][2 * B + u];     // (TB x D independent chains per stage: no partial sums needed)
#pragma unroll
                    for (int o = 1; o < W; ++o) mx += coef(r0, o) * xw[d][2 * B + u + o];                        // likelihoods.jl:129
                    const double e = M::f(d, xa, th) - mx;                                                       // :130
                    ew[d][2 * B + u] = (t1 >= 0 && t1 < n) ? e : 0.0;
                }
            }
            // ---- stage 2: t2 = s - 2B ----
#pragma unroll
            for (int u = 0; u < TB; ++u) {
                const double* cfs = cfb + u * CS;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const double2* r2 = reinterpret_cast<const double2*>(cfs + (2 * D + d) * WP);
                    double ke = coef(r2, 0) * ew[d][u];
#pragma unroll
                    for (int o = 1; o < W; ++o) ke += coef(r2, o) * ew[d][u + o];                                // :132 (1/beta1 in the table)
                    kw[d][2 * B + u] = ke;
                    eke[d] += ew[d][B + u] * ke;                                                                 // :146
                }
            }
            // ---- stage 3: t3 = s - 3B (outside the time axis x, Ke, the coefficients and hence every term are zero) ----
            double gvo[TB][D];
#pragma unroll
            for (int u = 0; u < TB; ++u) {
                const double* cfs = cfb + u * CS;
                double xa[D], wa[D];
#pragma unroll
                for (int d = 0; d < D; ++d) { xa[d] = xw[d][B + u]; wa[d] = kw[d][B + u]; }
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const double2* r1 = reinterpret_cast<const double2*>(cfs + (1 * D + d) * WP);
                    const double2* r3 = reinterpret_cast<const double2*>(cfs + (3 * D + d) * WP);
                    double cx = coef(r1, 0) * xw[d][u], mt = coef(r3, 0) * kw[d][u];
#pragma unroll
                    for (int o = 1; o < W; ++o) { cx += coef(r1, o) * xw[d][u + o]; mt += coef(r3, o) * kw[d][u + o]; }   // :133 (1/beta2 in the table), :192
                    const double2 yw = reinterpret_cast<const double2*>(cfs + 4 * D * WP)[d];    // (y, 1) for an observation, (0, 0) otherwise
                    const double e0 = (xa[d] - yw.x) * yw.y;                                                     // :123 (missing: e0 = 0)
                    double gv = -(e0 * obs_scale[d]);                                                            // :179
                    gv -= cx;                                                                                    // :186
                    gv += mt;                                                                                    // :194
                    M::jx_col_sub(d, xa, th, wa, gv);                                                            // :214-216
                    M::jth_row_sub(d, xa, th, wa[d], gth);                                                       // :219-221
                    bad |= want_grad && !isfinite(gv);
                    xcx[d] += xa[d] * cx;                                                                        // :150
                    sse[d] += e0 * e0;                                                                           // :139
                    gvo[u][d] = gv;
                }
            }
#pragma unroll
            for (int u = 0; u < TB; ++u)
#pragma unroll
                for (int d = 0; d < D; ++d) gs[(d * TCH + sb * TB + u) * CP + tid] = gvo[u][d];
            // the windows move TB places
#pragma unroll
            for (int d = 0; d < D; ++d) {
#pragma unroll
                for (int i = 0; i < XW - TB; ++i) xw[d][i] = xw[d][i + TB];
#pragma unroll
                for (int i = 0; i < EW - TB; ++i) { ew[d][i] = ew[d][i + TB]; kw[d][i] = kw[d][i + TB]; }
            }
        }
        __syncthreads();                     // WS: chunk ch + 1 has arrived, the staging warp may take this chunk's gradient tile
        if constexpr (!WS) { if (want_grad) stage_out(ch); }
    }
    __syncthreads();       // every x-gradient store of the block is issued before a guard below overwrites a chain's row

    // ---------------- per chain: log density in the reference's order of accumulation, sigma gradient, guards ----------------
    if (!valid) return;
    double* gp = want_grad ? a.grad + c * a.pitch : nullptr;
    const int P = a.P;
    if (a.sigma_invalid) {                                            // interface.jl:192-195
        a.ll[c] = -INFINITY;
        if (gp) for (int i = 0; i < P; ++i) gp[i] = NAN;
        return;
    }
    double ll = 0.0, prior = 0.0, gsig[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const double s = sig[d], s2 = s * s;
        const int nobs = a.nobs[d];
        double ll_obs = -0.5 * sse[d] / s2;                           // likelihoods.jl:139
        if (nobs > 0) ll_obs -= 0.5 * nobs * log(2.0 * M_PI * s2);   // :141
        ll += ll_obs / a.beta3;                                       // :143
        ll += -0.5 * eke[d];                                          // :146-147 (1/beta1 folded into K~)
        ll += -0.5 * xcx[d];                                          // :150-151 (1/beta2 folded into C~)
        gsig[d] = (s > 0 && nobs > 0) ? (sse[d] / s2 - nobs) / (s * a.beta3) : 0.0;   // :229-246
        if (gp) bad |= !isfinite(gsig[d]);
        if (!a.sigma_is_fixed) {
            const double raw = xp[nxt + d];
            prior += isnan(raw) ? raw : fmin(fmax(raw, -15.0), 15.0);  // interface.jl:206
        }
    }
    if (gp) {
#pragma unroll
        for (int i = 0; i < K; ++i) bad |= !isfinite(gth[i]);
    }
    bad |= !isfinite(ll);
    if (bad) {                                                        // interface.jl:222-226
        a.ll[c] = -INFINITY;
        if (gp) for (int i = 0; i < P; ++i) gp[i] = 0.0;
        return;
    }
    double total = ll;
    bool bad2 = false;
    double gls[D];
    if (!a.sigma_is_fixed) {
        total += prior;                                               // interface.jl:238
#pragma unroll
        for (int d = 0; d < D; ++d) { gls[d] = gsig[d] * sig[d] + 1.0; bad2 |= !isfinite(gls[d]); }   // :249-253
    }
    a.ll[c] = total;
    if (gp) {
        if (bad2) { for (int i = 0; i < P; ++i) gp[i] = 0.0; }       // interface.jl:260-264
        else {
#pragma unroll
            for (int i = 0; i < K; ++i) gp[n * D + i] = gth[i];
            if (!a.sigma_is_fixed) {
#pragma unroll
                for (int d = 0; d < D; ++d) gp[nxt + d] = gls[d];
            }
        }
    }
}

// Chains per block: the multiple of 32 (at most max_chains) that needs the fewest waves of slots = blocks-per-SM x sm_count blocks, the
// larger one on a tie.
inline int narrow_block_chains(int n_chains, long long slots, int max_chains) {
    int best = 32; long long best_cost = -1;
    for (int nc = 32; nc <= max_chains; nc += 32) {
        const long long tiles = (n_chains + nc - 1) / nc, waves = (tiles + slots - 1) / slots;
        const long long cost = waves * (100 + nc);          // a wave of wider blocks takes longer, but far less than proportionally
        if (best_cost < 0 || cost <= best_cost) { best = nc; best_cost = cost; }
    }
    return best;
}

template <int MODEL, int B, bool WS>
static cudaError_t narrow_launch_ws(const NarrowArgs& a, int sm_count, cudaStream_t st) {
    constexpr int TB = (B <= 2) ? 4 : 2, TCH = WS ? 8 : 16;
    constexpr int D = Ode<MODEL>::D;
    auto kern = narrow_logpost_kernel<MODEL, B, TB, TCH, WS>;
    const int nc = WS ? narrow_block_chains(a.n_chains, 2LL * sm_count, 96) : narrow_block_chains(a.n_chains, sm_count, kNarrowMaxChains);
    const int CP = nc + 1;
    const size_t smem = sizeof(double) * ((size_t)(WS ? 4 : 3) * D * TCH * CP + (size_t)2 * TCH * a.CS);
    static PerDeviceOnce once;
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WS ? 100 * 1024 : 227 * 1024);
        if (e != cudaSuccess) return e;
    }
    const int blocks = (a.n_chains + nc - 1) / nc;
    kern<<<blocks, nc + (WS ? 32 : 0), smem, st>>>(a);
    return cudaGetLastError();
}

// Which organisation (measured, FN n=201, one B200, ms at 8192 / 16 384 / 65 536 chains; WS = false | true):
//   b = 1:  0.043 / 0.068 / 0.187  |  0.037 / 0.068 / 0.199        b = 2:  0.051 / 0.076 / 0.217  |  0.039 / 0.072 / 0.248
//   b = 4:  0.084 / 0.091 / 0.306  |  0.070 / 0.113 / 0.280
// Other tries, slower: four 4-warp blocks per SM at 128 registers (b = 1: 0.250), two chains per thread sharing the coefficient loads
// (b = 1: 0.281 ms, with half the warps), 7 + 1 warps (0.30 ms at every b).
template <int MODEL, int B>
static cudaError_t narrow_launch_b(const NarrowArgs& a, int sm_count, cudaStream_t st) {
    const long long per_sm = (a.n_chains + sm_count - 1) / sm_count;
    // chains per SM: the warp-specialised kernel up to its own first wave (2 x 96; b >= 3: while 64-chain blocks do), and for b >= 3
    // again beyond one wave of the other (b = 4: 32 768 chains 0.153 | 0.211 ms, 65 536 chains 0.306 | 0.281)
    const bool ws = (B >= 3) ? (per_sm <= 64 || per_sm > 256) : per_sm <= 96;
    return ws ? narrow_launch_ws<MODEL, B, true>(a, sm_count, st) : narrow_launch_ws<MODEL, B, false>(a, sm_count, st);
}

template <int MODEL>
cudaError_t narrow_launch_model(const NarrowArgs& a, int b, int sm_count, cudaStream_t st) {
    switch (b) {
    case 0: return narrow_launch_b<MODEL, 0>(a, sm_count, st);
    case 1: return narrow_launch_b<MODEL, 1>(a, sm_count, st);
    case 2: return narrow_launch_b<MODEL, 2>(a, sm_count, st);
    case 3: return narrow_launch_b<MODEL, 3>(a, sm_count, st);
    case 4: return narrow_launch_b<MODEL, 4>(a, sm_count, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace magi
