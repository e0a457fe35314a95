// K1 "flow": fused banded log-posterior + gradient as a DATAFLOW of symmetric warps (FP64 tensor cores, DMMA.8x8x4).
//
// Replaces, for a batch of independent chains, the reference's
//   log_likelihood_and_gradient_banded          src/likelihoods.jl:43-257
//   LogDensityProblems.logdensity_and_gradient  src/logdensityproblems_interface.jl:176-267
//
// Same tensor formulation as banded_kernel.cuh (8 chains = M of a DMMA, 8 output times = N, contraction over 4-time
// chunks, NCH = 2 HB + 2 chunks per output tile), different organisation -- chosen after the measurements of round 1
// (profiles/README.md: the warp-specialised kernel spends its time in queue / ring hand-offs and window moves, 36 % of
// the issue slots used, DMMA pipe 47 % busy):
//   * one block = 8 G chains (G = 2) and ALL of their state in shared memory: X, E = f - m~x and KE = K~e as
//     [chain][dim][time] rows, zero-padded by 4 HB on both sides (no bounds predicates in the product loops) and
//     XOR-swizzled per chain row (A-fragment LDS.64 and the 128-bit pointwise accesses are both conflict-free);
//   * the work is cut into UNITS (sweep, dimension, pair of output tiles):
//        S1  mx = m~ x;  e = f(x, theta) - mx -> E                                   (likelihoods.jl:129-130)
//        S2  Ke = K~ e -> KE;  sum e.Ke                                              (:132, :146)
//        S3  Cx = C~ x, mt = m~^T Ke;  gradient incl. the ODE Jacobian terms; sums   (:133, :150, :179-221)
//     16 identical warps draw units from a ticket counter in a fixed wavefront order (S1 runs HP + 1 pairs ahead of S2,
//     S2 HP + 1 pairs ahead of S3); a unit waits for the units it reads from on per-unit mbarriers.  No warp roles, no
//     queues, no rings: whenever one warp does pointwise work or waits, the other three of its SM sub-partition issue DMMAs;
//   * the A operand (state) comes from shared memory (14 LDS.64 per chain group and product of a unit), the B operand
//     (band-table fragments, `fragtab` in "natural" order) goes from L2 STRAIGHT INTO REGISTERS with one coalesced
//     16-byte load per lane and chunk, and is used for both chain groups and both tiles of the pair (4 DMMAs per load);
//     nothing is staged, so no register window slides (round 1: 1.5 IMAD.MOV per DMMA);
//   * reductions go to per-unit slots and are summed in a fixed order by the final per-chain stage: results do not
//     depend on which warp ran which unit (bit-reproducible).
// The block is persistent over chain blocks (grid = min(blocks, SMs)); mbarrier phases alternate per chain block.
// Used when the state of 16 chains fits shared memory (flow_smem_bytes); other shapes run banded_kernel.cuh.
#pragma once
#include <cmath>
#include <type_traits>
#include "magi_common.cuh"
#include "k1_primitives.cuh"
#include "ode_models.cuh"

namespace magi {

__device__ __forceinline__ double2 ldg_frag(const double2* p) {
    double2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void unit_done(unsigned long long* bar) {      // lane 0, after __syncwarp(): the unit's shared-memory stores are visible
#ifdef MAGI_FLOW_RELAXED_ARRIVE
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(bar)) : "memory");
#else
    mbar_arrive(bar);
#endif
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

#ifdef MAGI_FLOW_TIMELINE
#define FLOW_TL(slot, t_from) do { const long long now_ = clock64(); dbg_t[slot] += now_ - t_from; t_from = now_; } while (0)
#else
#define FLOW_TL(slot, t_from) do { } while (0)
#endif

#ifndef MAGI_FLOW_WARPS
#define MAGI_FLOW_WARPS 16
#endif
constexpr int kFlowWarps = MAGI_FLOW_WARPS;

template <int MODEL, int HB, int G>
__global__ void __launch_bounds__(32 * kFlowWarps, 1) flow_logpost_kernel(const FlowArgs a) {
    using M = Ode<MODEL>;
    constexpr int D = M::D, K = M::K, KX = M::KX;
    constexpr int NCH = 2 * HB + 2, CH = 8 * G, RED = 3 + K;
    constexpr int HP = (4 * HB + 15) / 16;               // pairs of halo on each side of an output pair
    constexpr int PCW = KX + 3 * D;                      // per-chain constants: theta + invariants | 1/(sigma^2 beta3) | sigma | clamped log sigma
    // depth of the fragment ring (16-byte loads in flight per lane): a divisor of NCH, so that chunk j of every product sits in slot j % FD
    constexpr int FD = NCH <= 8 ? NCH : (NCH % 6 == 0 ? 6 : (NCH % 8 == 0 ? 8 : (NCH % 7 == 0 ? 7 : 5)));
    static_assert(NCH % FD == 0, "fragment ring depth must divide the chunk count");
    struct Frags { double2 v[FD]; };                     // chunks hh .. hh + FD - 1 of the running product: tiles 2p (x) and 2p + 1 (y)
    extern __shared__ __align__(128) double smem[];
    const int n = a.n, NP = a.NP, RS0 = a.RS0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
    const size_t arr = (size_t)CH * D * RS0;
    double* B0 = smem;                                   // [CH][D][RS0]  position of time t in a row: (t + 4 HB) ^ sw(chain)
    double* B1 = B0 + arr;                               // X and E alternate between B0 and B1 from chain block to chain block
    double* Ks = B1 + arr;
    double* red = Ks + arr;                              // [D][NP][CH][RED]: e.Ke | x.Cx | sse | theta-gradient partials
    double* PC = red + (size_t)D * NP * CH * RED;        // [2][CH][PCW], double-buffered over chain blocks
    double* FIN = PC + 2 * CH * PCW;                     // [CH][D][RED] sums over the tile pairs, then [CH][D][4] terms per dimension
    double* Ys = FIN + CH * D * (RED + 4);                      // [D][16 NP] observations (NaN beyond n)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(Ys + (size_t)D * 16 * NP);   // [2][D][NP]: S1 done, S2 done
    int* ticket = reinterpret_cast<int*>(bars + 2 * D * NP);     // [0] unit ticket, [1] row ticket of the next chain block's state load
    int* units = ticket + 4;                             // [n_units] the unit order (read once per unit: keep it out of L2 latency)

    // ---- once per block: barriers, zero X / E / KE (their padding is never written afterwards), observations ----
    for (int i = threadIdx.x; i < 2 * D * NP; i += blockDim.x) mbar_init(bars + i, 1);
    for (size_t i = threadIdx.x; i < 3 * arr; i += blockDim.x) B0[i] = 0.0;
    for (int i = threadIdx.x; i < D * 16 * NP; i += blockDim.x) {
        const int dd = i / (16 * NP), t = i % (16 * NP);
        Ys[i] = (t < n) ? a.yobs[(size_t)dd * n + t] : __longlong_as_double(0x7ff8000000000000LL);
    }
    for (int i = threadIdx.x; i < a.n_units; i += blockDim.x) units[i] = a.units[i];
    if (threadIdx.x == 0) { ticket[0] = 0; ticket[1] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    __syncthreads();

    const int sw = ((gid & 1) << 3) | ((gid & 2) << 1);  // row swizzle: chains 0..3 (mod 4) -> 0, 8, 4, 12
    int crow[G];                                         // offset of this lane's chain row (dimension 0) for each chain group
#pragma unroll
    for (int g = 0; g < G; ++g) crow[g] = ((g * 8 + gid) * D) * RS0;
    const double inv_b3 = a.inv_beta[2];
    const double2* ftab = reinterpret_cast<const double2*>(a.fragtab);
    auto frag_ptr = [&](int view, int d, int p) { return ftab + ((size_t)((view * D + d) * NP + p) * NCH) * 32 + lane; };

    // acc[g][tt] += sum_hh  S[chain group g][chunk 4p + 2tt + hh]  x  frag[hh].{x: tile 2p, y: tile 2p + 1}.  The fragments
    // come through a ring of FD registers per lane: slot hh % FD holds chunk hh on entry for hh < FD (prefetched by the
    // previous product of this warp) and is refilled, right after its use, with chunk hh + FD of this product or -- for
    // the last FD chunks -- with chunk hh + FD - NCH of the NEXT product this warp will run (`next`, null: none): the L2
    // latency of the fragments hides behind the DMMAs.
    auto band_product = [&](const double* S, int d, int p, Frags& fb, double (&acc)[G][2][2], const double2* cur, const double2* next) {
        double av[G][NCH + 2];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const double* row = S + crow[g] + d * RS0 + 16 * p + q;
            const double* ab[4] = {row + (0 ^ sw), row + (4 ^ sw), row + (8 ^ sw), row + (12 ^ sw)};
#pragma unroll
            for (int k = 0; k < NCH + 2; ++k) av[g][k] = ab[k & 3][16 * (k >> 2)];
        }
#pragma unroll
        for (int hh = 0; hh < NCH; ++hh) {
            const double2 f = fb.v[hh % FD];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                dmma884(acc[g][0][0], acc[g][0][1], av[g][hh], f.x);
                dmma884(acc[g][1][0], acc[g][1][1], av[g][hh + 2], f.y);
            }
            if (hh + FD < NCH) fb.v[hh % FD] = ldg_frag(cur + (hh + FD) * 32);
            else if (next != nullptr) fb.v[hh % FD] = ldg_frag(next + (hh + FD - NCH) * 32);
        }
    };
    // position (before the row offset) of the two consecutive times 16p + 8tt + 2q (+1) this lane owns in a C fragment
    auto cpos = [&](int p, int tt) { return (4 * HB + 16 * p + 8 * tt + 2 * q) ^ sw; };

    // State of chain block cb -> X buffer `dst`: one (chain, dimension) row per call, 8-byte cp.async (rows are only 8-byte
    // aligned: P is odd in general).  Row index CH * D stands for the per-chain constants (theta + invariants, sigma terms).
    auto load_row = [&](int cb, int buf, double* dst, int r) {
        const long long chain0 = (long long)cb * CH;
        if (r < CH * D) {
            const int c = r / D, dd = r % D;
            long long ch = chain0 + c; if (ch >= a.n_chains) ch = a.n_chains - 1;
            const double* src = a.params + ch * a.pitch + (size_t)dd * n;
            double* row = dst + (size_t)r * RS0;
            const int s = ((c & 1) << 3) | ((c & 2) << 1);
            for (int t = lane; t < n; t += 32) cp_async8(row + ((t + 4 * HB) ^ s), src + t);
        } else if (lane < CH) {
            const int c = lane;
            long long ch = chain0 + c; if (ch >= a.n_chains) ch = a.n_chains - 1;
            const double* cp = a.params + ch * a.pitch + (size_t)n * D;
            double* pc = PC + ((size_t)buf * CH + c) * PCW;
            double th[KX];
#pragma unroll
            for (int i = 0; i < K; ++i) th[i] = cp[i];
            M::prepare(th);
#pragma unroll
            for (int i = 0; i < KX; ++i) pc[i] = th[i];
#pragma unroll
            for (int dd = 0; dd < D; ++dd) {
                double s, cl = 0.0;
                if (a.sigma_is_fixed) s = a.sigma_init[dd];
                else {
                    const double raw = cp[K + dd];
                    cl = isnan(raw) ? raw : fmin(fmax(raw, -15.0), 15.0);               // interface.jl:200
                    s = isnan(raw) ? raw : exp(cl);
                }
                pc[KX + dd] = (1.0 / (s * s)) * inv_b3;
                pc[KX + D + dd] = s;
                pc[KX + 2 * D + dd] = cl;
            }
        }
    };
    static_assert(CH <= 32, "the per-chain constants are computed by one warp");

#ifdef MAGI_FLOW_TIMELINE
    long long dbg_t[16];                                 // instrumented builds only (tools/build_variant.py tl -DMAGI_FLOW_TIMELINE):
#pragma unroll                                           // load wait, unit loop, dependency waits, loop tail, final, units, per-sweep timeline
    for (int i = 0; i < 16; ++i) dbg_t[i] = 0;
#define FLOW_DBG(stmt) do { if (a.dbg) { stmt; } } while (0)
#else
#define FLOW_DBG(stmt) do { } while (0)
#endif
    if (blockIdx.x < a.n_cblocks)
        for (int r = warp; r <= CH * D; r += kFlowWarps) load_row(blockIdx.x, 0, B0, r);

    for (int cb = blockIdx.x, it = 0; cb < a.n_cblocks; cb += gridDim.x, ++it) {
        const unsigned par = it & 1;
        const long long chain0 = (long long)cb * CH;
        const double* pcb = PC + (size_t)(it & 1) * CH * PCW;
        double* Xc = (it & 1) ? B1 : B0;                     // this chain block's X; E goes to the other buffer, which then
        double* Ec = (it & 1) ? B0 : B1;                     // receives the NEXT chain block's X once every S2 unit is done
        [[maybe_unused]] long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0;
        FLOW_DBG(tk0 = clock64());
        cp_async_wait_all();
        __syncthreads();                                     // X and the per-chain constants of this chain block are in place
        FLOW_DBG(tk1 = clock64());
        auto dep_wait = [&](unsigned long long* bar) {
#ifdef MAGI_FLOW_TIMELINE
            if (a.dbg) { const long long w0 = clock64(); mbar_wait(bar, par); dbg_t[2] += clock64() - w0; return; }
#endif
            mbar_wait(bar, par);
        };

        // ================= units (fb holds the fragments of the unit's first product on entry) =================
        auto s1_unit = [&](auto dconst, int p, Frags& fb, const double2* next) {   // mx = m~ x_d (likelihoods.jl:129); e = f_d(x, theta) - mx (:130)
            constexpr int d = decltype(dconst)::value;
            [[maybe_unused]] long long tl = 0;
#ifdef MAGI_FLOW_TIMELINE
            tl = clock64();
#endif
            double acc[G][2][2];
#pragma unroll
            for (int g = 0; g < G; ++g) { acc[g][0][0] = acc[g][0][1] = acc[g][1][0] = acc[g][1][1] = 0.0; }
            band_product(Xc, d, p, fb, acc, frag_ptr(0, d, p), next);
            FLOW_TL(7, tl);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                double th[KX];
#pragma unroll
                for (int i = 0; i < KX; ++i) th[i] = pcb[(g * 8 + gid) * PCW + i];
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) {
                    const int pos = cpos(p, tt), t0 = 16 * p + 8 * tt + 2 * q;
                    double x0[D], x1[D];
#pragma unroll
                    for (int dd = 0; dd < D; ++dd) {
                        const double2 xv = *reinterpret_cast<const double2*>(Xc + crow[g] + dd * RS0 + pos);
                        x0[dd] = xv.x; x1[dd] = xv.y;
                    }
                    double e0 = M::f(d, x0, th) - acc[g][tt][0];
                    double e1 = M::f(d, x1, th) - acc[g][tt][1];
                    if (t0 >= n) e0 = 0.0;
                    if (t0 + 1 >= n) e1 = 0.0;
                    *reinterpret_cast<double2*>(Ec + crow[g] + d * RS0 + pos) = make_double2(e0, e1);
                }
            }
            __syncwarp();
            FLOW_TL(8, tl);
            if (lane == 0) unit_done(bars + (0 * D + d) * NP + p);
            FLOW_TL(6, tl);
        };

        auto s2_unit = [&](int d, int p, Frags& fb, const double2* next) {          // Ke = K~ e_d / beta1 (likelihoods.jl:132), e.Ke (:146)
            [[maybe_unused]] long long tl = 0;
#ifdef MAGI_FLOW_TIMELINE
            tl = clock64();
#endif
            {
                const int j0 = max(p - HP, 0), j1 = min(p + HP, NP - 1);
                for (int j = j0; j <= j1; ++j) dep_wait(bars + (0 * D + d) * NP + j);
            }
            FLOW_TL(9, tl);
            double acc[G][2][2];
#pragma unroll
            for (int g = 0; g < G; ++g) { acc[g][0][0] = acc[g][0][1] = acc[g][1][0] = acc[g][1][1] = 0.0; }
            band_product(Ec, d, p, fb, acc, frag_ptr(2, d, p), next);
            FLOW_TL(10, tl);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                double eke = 0.0;
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) {
                    const int pos = cpos(p, tt);
                    const double2 ev = *reinterpret_cast<const double2*>(Ec + crow[g] + d * RS0 + pos);
                    eke += ev.x * acc[g][tt][0];
                    eke += ev.y * acc[g][tt][1];
                    *reinterpret_cast<double2*>(Ks + crow[g] + d * RS0 + pos) = make_double2(acc[g][tt][0], acc[g][tt][1]);
                }
                eke = quad_sum(eke);
                if (q == 0) red[((size_t)(d * NP + p) * CH + g * 8 + gid) * RED + 0] = eke;
            }
            __syncwarp();
            FLOW_TL(11, tl);
            if (lane == 0) unit_done(bars + (1 * D + d) * NP + p);
            FLOW_TL(6, tl);
        };

        auto s3_unit = [&](auto dconst, int p, Frags& fb, const double2* next) {   // Cx = C~ x_d / beta2 (:133), mt = m~^T Ke_d (:192), gradient (:179-221)
            constexpr int d = decltype(dconst)::value;
            [[maybe_unused]] long long tl = 0;
#ifdef MAGI_FLOW_TIMELINE
            tl = clock64();
#endif
            double c[G][2][2], um[G][2][2];
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) { c[g][tt][0] = c[g][tt][1] = 0.0; um[g][tt][0] = um[g][tt][1] = 0.0; }
            band_product(Xc, d, p, fb, c, frag_ptr(1, d, p), frag_ptr(3, d, p));          // (no dependency: runs while S2 units may still be in flight)
            FLOW_TL(12, tl);
            {
                const int j0 = max(p - HP, 0), j1 = min(p + HP, NP - 1);
                for (int j = j0; j <= j1; ++j) dep_wait(bars + (1 * D + d) * NP + j);
#pragma unroll
                for (int dd = 0; dd < D; ++dd)
                    if (dd != d) dep_wait(bars + (1 * D + dd) * NP + p);
            }
            FLOW_TL(13, tl);
            band_product(Ks, d, p, fb, um, frag_ptr(3, d, p), next);
            FLOW_TL(14, tl);
            const double* yd = Ys + d * 16 * NP;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                double th[KX];
#pragma unroll
                for (int i = 0; i < KX; ++i) th[i] = pcb[(g * 8 + gid) * PCW + i];
                const double obs_scale = pcb[(g * 8 + gid) * PCW + KX + d];
                const long long ch = chain0 + g * 8 + gid;
                double* gout = (a.grad != nullptr && ch < a.n_chains) ? a.grad + ch * a.pitch + (size_t)d * n : nullptr;
                double xcx = 0.0, sse = 0.0, gth[K];
                bool bad = false;
#pragma unroll
                for (int i = 0; i < K; ++i) gth[i] = 0.0;
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) {
                    const int pos = cpos(p, tt), t0 = 16 * p + 8 * tt + 2 * q;
                    double xv[2][D], w[2][D];
#pragma unroll
                    for (int dd = 0; dd < D; ++dd) {
                        const double2 x2 = *reinterpret_cast<const double2*>(Xc + crow[g] + dd * RS0 + pos);
                        const double2 k2 = *reinterpret_cast<const double2*>(Ks + crow[g] + dd * RS0 + pos);     // already Ke / beta1 (:201)
                        xv[0][dd] = x2.x; xv[1][dd] = x2.y;
                        w[0][dd] = k2.x; w[1][dd] = k2.y;
                    }
                    const double2 y2 = *reinterpret_cast<const double2*>(yd + t0);
                    const double yy[2] = {y2.x, y2.y};
#pragma unroll
                    for (int pt = 0; pt < 2; ++pt) {
                        const bool valid = t0 + pt < n;
                        const double xdv = xv[pt][d];
                        const bool fin = valid && isfinite(yy[pt]);                      // likelihoods.jl:123
                        const double e0 = fin ? xdv - yy[pt] : 0.0;
                        double gv = (um[g][tt][pt] - c[g][tt][pt]) - e0 * obs_scale;     // :179, :186, :194 (1/beta folded into the tables)
                        M::jx_col_sub(d, xv[pt], th, w[pt], gv);                         // :214-216
                        M::jth_row_sub(d, xv[pt], th, valid ? w[pt][d] : 0.0, gth);      // :219-221
                        sse += e0 * e0;                                                  // :139, :234
                        xcx += xdv * c[g][tt][pt];                                       // :150
                        bad |= valid && !isfinite(gv);
                        if (valid && gout != nullptr) gout[t0 + pt] = gv;
                    }
                }
                xcx = quad_sum(xcx); sse = quad_sum(sse);
#pragma unroll
                for (int i = 0; i < K; ++i) gth[i] = quad_sum(gth[i]);
                const unsigned badm = __ballot_sync(0xffffffffu, bad);
                if (q == 0) {
                    double* r = red + ((size_t)(d * NP + p) * CH + g * 8 + gid) * RED;
                    // a non-finite gradient entry poisons the chain (interface.jl:222-226); value-only calls look at ll alone (:155-160)
                    const bool poison = ((badm >> (gid * 4)) & 0xfu) && a.grad != nullptr;
                    r[1] = xcx;
                    r[2] = poison ? __longlong_as_double(0x7ff8000000000000LL) : sse;
#pragma unroll
                    for (int i = 0; i < K; ++i) r[3 + i] = gth[i];
                }
            }
            FLOW_TL(15, tl);
        };

        // ---- ticket loop: every warp draws the next unit of the list, one unit ahead of the one it runs (so that it can
        //      prefetch that unit's fragments during its own last product) ----
        auto take_ticket = [&]() {
            int u = 0;
            if (lane == 0) u = atomicAdd(ticket, 1);      // ticket[0]
            return __shfl_sync(0xffffffffu, u, 0);
        };
        auto first_ptr = [&](int u) -> const double2* {       // fragments of the first product of unit u (S1: m~, S2: K~, S3: C~)
            if (u >= a.n_units) return nullptr;
            const int code = units[u];
            const int s = code & 3, d = (code >> 2) & 63, p = code >> 8;
            return frag_ptr(s == 0 ? 0 : (s == 1 ? 2 : 1), d, p);
        };
        {
            int u = take_ticket();
            Frags fb;
            if (u < a.n_units) {
                const double2* fr = first_ptr(u);
#pragma unroll
                for (int hh = 0; hh < FD; ++hh) fb.v[hh] = ldg_frag(fr + hh * 32);
            }
            while (u < a.n_units) {
                const int un = take_ticket();
                const double2* next = first_ptr(un);
                FLOW_DBG(dbg_t[5] += 1);
                const int code = units[u];
                const int s = code & 3, d = (code >> 2) & 63, p = code >> 8;
                if (s == 0) { auto f = [&](auto dc) { s1_unit(dc, p, fb, next); }; dispatch_dim<D>(d, f); }
                else if (s == 1) s2_unit(d, p, fb, next);
                else { auto f = [&](auto dc) { s3_unit(dc, p, fb, next); }; dispatch_dim<D>(d, f); }
                u = un;
            }
        }
        FLOW_DBG(tk2 = clock64());
        // A warp that has run out of units starts the NEXT chain block's state load into the E buffer (free once every S2
        // unit is done; the S3 units still running read X and KE only): rows are drawn from a second ticket counter, so the
        // first warps to finish issue the whole load and its HBM latency overlaps the tail of the unit loop.
        if (cb + (int)gridDim.x < a.n_cblocks) {
            for (int i = lane; i < D * NP; i += 32) mbar_wait(bars + D * NP + i, par);
            __syncwarp();
            for (;;) {
                int r = 0;
                if (lane == 0) r = atomicAdd(ticket + 1, 1);
                r = __shfl_sync(0xffffffffu, r, 0);
                if (r > CH * D) break;
                load_row(cb + gridDim.x, (it + 1) & 1, Ec, r);
            }
        }
        __syncthreads();                                     // every unit of this chain block is done
        FLOW_DBG(tk3 = clock64());
        if (threadIdx.x == 0) { ticket[0] = 0; ticket[1] = 0; }

        // ---------------- final (warps 0 .. NWF - 1; the others go on to wait for the next chain block's state):
        //   A  one thread per (chain, dimension, quantity): sum of the per-unit partials over the tile pairs, in a fixed order
        //   B  one thread per (chain, dimension): the terms of that dimension (log, divisions)
        //   C  one thread per chain: accumulation in the reference's order, guards, theta / sigma gradient
        constexpr int NSUM = CH * D * RED, NWF = (NSUM + 31) / 32;
        if (warp < NWF) {
            if (threadIdx.x < NSUM) {
                const int i = threadIdx.x % RED, cd = threadIdx.x / RED, d = cd % D, c = cd / D;
                const double* r = red + ((size_t)(d * NP) * CH + c) * RED + i;
                double v = 0.0;
                for (int p = 0; p < NP; ++p) v += r[(size_t)p * CH * RED];
                FIN[threadIdx.x] = v;                                 // [c][d][RED]
            }
            named_barrier(1, NWF * 32);
            if (threadIdx.x < CH * D) {
                const int c = threadIdx.x / D, d = threadIdx.x % D;
                const double* v = FIN + (size_t)threadIdx.x * RED;
                const double s = pcb[c * PCW + KX + D + d], s2 = s * s;
                const int nobs = a.nobs[d];
                double ll_obs = -0.5 * v[2] / s2;                     // likelihoods.jl:139
                if (nobs > 0) ll_obs -= 0.5 * nobs * log(2.0 * M_PI * s2);   // :141
                double* f = FIN + NSUM + (size_t)threadIdx.x * 4;
                f[0] = ll_obs / a.beta[2];                            // :143
                f[1] = -0.5 * v[0];                                   // :146-147 (1/beta1 folded into K~)
                f[2] = -0.5 * v[1];                                   // :150-151 (1/beta2 folded into C~)
                f[3] = (s > 0 && nobs > 0) ? (v[2] / s2 - nobs) / (s * a.beta[2]) : 0.0;   // :229-246
            }
            named_barrier(1, NWF * 32);
            if (threadIdx.x < CH) {
                const long long c = chain0 + threadIdx.x;
                if (c < a.n_chains) {
                    double* gp = a.grad ? a.grad + c * a.pitch : nullptr;
                    const int nxt = n * D + K, P = a.P;
                    if (a.sigma_invalid) {                                    // interface.jl:192-195
                        a.ll[c] = -INFINITY;
                        if (gp) for (int i = 0; i < P; ++i) gp[i] = NAN;
                    } else {
                        double ll = 0.0, prior = 0.0;
                        double gsig[D], sig[D], gthf[K];
                        bool bad = false;
#pragma unroll
                        for (int i = 0; i < K; ++i) gthf[i] = 0.0;
#pragma unroll
                        for (int d = 0; d < D; ++d) {                         // the reference's order of accumulation
                            const double* f = FIN + NSUM + ((size_t)threadIdx.x * D + d) * 4;
                            const double* v = FIN + ((size_t)threadIdx.x * D + d) * RED;
                            ll += f[0]; ll += f[1]; ll += f[2];
                            gsig[d] = f[3];
#pragma unroll
                            for (int i = 0; i < K; ++i) gthf[i] += v[3 + i];
                            sig[d] = pcb[threadIdx.x * PCW + KX + D + d];
                            prior += pcb[threadIdx.x * PCW + KX + 2 * D + d];   // interface.jl:206 (0 when sigma is fixed)
                            if (gp) bad |= !isfinite(gsig[d]);
                        }
                        if (gp) {
#pragma unroll
                            for (int i = 0; i < K; ++i) bad |= !isfinite(gthf[i]);
                        }
                        bad |= !isfinite(ll);
                        if (bad) {                                            // interface.jl:222-226 (value only: :155-160)
                            a.ll[c] = -INFINITY;
                            if (gp) for (int i = 0; i < P; ++i) gp[i] = 0.0;
                        } else {
                            double total = ll;
                            bool bad2 = false;
                            double gls[D];
                            if (!a.sigma_is_fixed) {
                                total += prior;                               // interface.jl:238
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
            }
        }
        FLOW_DBG(dbg_t[0] += tk1 - tk0; dbg_t[1] += tk2 - tk1; dbg_t[3] += tk3 - tk2; dbg_t[4] += clock64() - tk3);
    }
#ifdef MAGI_FLOW_TIMELINE
    if (a.dbg && lane == 0) {
        long long* o = a.dbg + ((size_t)blockIdx.x * kFlowWarps + warp) * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = dbg_t[i];
    }
#endif
}

template <int MODEL, int HB, int G>
static cudaError_t flow_launch_g(const FlowArgs& a, int grid, size_t smem_bytes, cudaStream_t st) {
    auto kern = flow_logpost_kernel<MODEL, HB, G>;
    static PerDeviceOnce once;       // per instantiation
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, 32 * kFlowWarps, smem_bytes, st>>>(a);
    return cudaGetLastError();
}

template <int MODEL, int HB>
static cudaError_t flow_launch_one(const FlowArgs& a, int grid, size_t smem_bytes, cudaStream_t st) {
    return a.G == 1 ? flow_launch_g<MODEL, HB, 1>(a, grid, smem_bytes, st) : flow_launch_g<MODEL, HB, 2>(a, grid, smem_bytes, st);
}

template <int MODEL>
cudaError_t flow_launch_model(const FlowArgs& a, int HB, int grid, size_t smem_bytes, cudaStream_t st) {
    switch (HB) {
    case 0: return flow_launch_one<MODEL, 0>(a, grid, smem_bytes, st);
    case 1: return flow_launch_one<MODEL, 1>(a, grid, smem_bytes, st);
    case 2: return flow_launch_one<MODEL, 2>(a, grid, smem_bytes, st);
    case 3: return flow_launch_one<MODEL, 3>(a, grid, smem_bytes, st);
    case 4: return flow_launch_one<MODEL, 4>(a, grid, smem_bytes, st);
    case 5: return flow_launch_one<MODEL, 5>(a, grid, smem_bytes, st);
    case 6: return flow_launch_one<MODEL, 6>(a, grid, smem_bytes, st);
    case 7: return flow_launch_one<MODEL, 7>(a, grid, smem_bytes, st);
    case 8: return flow_launch_one<MODEL, 8>(a, grid, smem_bytes, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace magi
