// K1: fused banded log-posterior + gradient on FP64 tensor cores (DMMA.8x8x4).
//
// Replaces, for a batch of independent chains, the reference's
//   log_likelihood_and_gradient_banded          src/likelihoods.jl:43-257
//   LogDensityProblems.logdensity_and_gradient  src/logdensityproblems_interface.jl:176-267
//
// Formulation.  For every dimension d the four band products (m~ x, K~ e, C~ x, m~^T Ke; likelihoods.jl:129,132,133,192)
// are written as (chains x time) = (chains x time) . (band table) products: 8 chains form the M extent of a DMMA.8x8x4,
// 8 output times its N extent, and the contraction runs over 4-time chunks.  The operand (x, e, Ke) lives in a register
// window that slides along the time axis, so every state value enters the tensor pipe from registers and the only
// per-DMMA load is the table fragment (staged once per block in shared memory by TMA).  The time->slot permutation
// (lane (gid,q) owns times 8J+q and 8J+q+4 of tile J) makes the C fragment of one product directly usable as the A
// fragment of the next.
//
//   phase A1:  mx = m~ x_d;  e = f_d(x, theta) - mx (pointwise warps);  Ke = K~ e  -> Ke scratch; sum e.Ke (DMMA warps)
//   phase A2:  Cx = C~ x_d;  mt = m~^T Ke_d (DMMA warps);  pointwise gradient incl. the ODE Jacobian terms, which need Ke of ALL
//              dimensions at the same time point (hence the block-wide barrier in between)
//   final (per chain):  log-likelihood assembly in the reference's term order, sigma gradient, log-sigma transform and the
//              per-chain -Inf / zero-gradient guards (interface.jl:179-264).
// Warp roles, exchange protocol and the measured behaviour: see the comment above the kernel and DESIGN.md section 4.
#pragma once
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include "magi_common.cuh"
#include "k1_primitives.cuh"
#include "ode_models.cuh"

namespace magi {

// Warp-specialised K1 (v13).  One block = G chain-groups (8 chains each) x D dimensions = G*D tasks; every task has a DMMA
// warp ("C") and a pointwise warp ("P"), 2*G*D <= 16 warps, <= 128 registers, one block per SM.  Operand windows (x, e, Ke)
// live in registers, are prefilled in a prologue (every step of a sweep produces an output pair) and slide two 8-time tiles
// per step.  All hand-offs are ONE-DIRECTIONAL queues of kXStages stages in
// shared memory (a "full" and an "empty" mbarrier per stage), so neither warp waits for a round trip through the other:
//   A1  P(u): mx = m~ x (x window), e = f(x, theta) - mx at tiles (Ja, Ja+1) -> queue (P runs ahead)
//       C(i): e window <- queue; Ke pair = K~ e -> Ke scratch, sum e.Ke            (four DMMA-issuing warps per sub-partition)
//   A2  C(u): Cx, m~^T Ke tiles (Jc, Jc+1) from the x and Ke windows, combined; x.Cx -> queue (C runs ahead); the two DMMA
//       warps of an SM sub-partition take turns on the tensor pipe (named-barrier ping-pong)
//       P(u): gradient incl. Jacobian terms, stores, sums (predicate-free variant for steps inside the time axis)
// Fragment pair-blocks arrive by TMA bulk copy into two kRingStages-deep rings per dimension (ring A: m~ for the P warps in A1,
// C~ in A2; ring B: K~ in A1, m~^T in A2, for the C warps): full[stage] (expect_tx) / empty[stage] (G arrivals) mbarriers,
// issued two uses ahead by an elected thread of the consumer warp whose turn it is (the duty rotates); one 16-byte LDS feeds
// the same chunk of both tiles of a pair, i.e. two independent accumulate chains.
// Measured behaviour, what limits it and every variant tried: DESIGN.md section 4, profiles/README.md.
constexpr int kXStages = 3, kRingStages = 3, kXSlots = 4;

template <int MODEL, int HB>
__global__ void __launch_bounds__(512, 1) banded_logpost_kernel(const BandedArgs a) {
    using M = Ode<MODEL>;
    constexpr int D = M::D, K = M::K;
    constexpr int NCH = 2 * HB + 2, LAGT = (HB + 1) / 2, W2 = 2 * LAGT + 4 + HB;
    constexpr int RED = 4 + K;   // e.Ke, x.Cx, sse, bad flag, theta-gradient partials
    constexpr int BLKP = NCH * 64;              // doubles per fragment pair-block (one view, two tiles interleaved per lane)
    // Every view is stored in pairs of tiles (2p, 2p + 1).  The products with m~, C~, m~^T run LAGT tiles behind the window head;
    // when LAGT is odd the window heads are the tile pairs (2u - 1, 2u) (HO = 1), so that the OUTPUT pairs are (2p, 2p + 1) as
    // well and no output pair straddles an end of the time axis; the e window is 2 HO chunks longer so that the K~ output pairs
    // (LAGT + HO tiles behind the e pairs) are aligned the same way.
    constexpr int HO = LAGT & 1, LAGH = (LAGT + HO) / 2, W2E = W2 + 2 * HO;
    constexpr int S = kXStages, R = kRingStages, XS = kXSlots;
    extern __shared__ __align__(128) double smem[];
    const int NT = a.NT, n = a.n, G = a.G;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
    const int ntask = G * D;
    const int role = warp / ntask;               // 0: DMMA warp (C), 1: pointwise warp (P)
    const int task = warp % ntask, g = task / D, d_rt = task % D;   // the two DMMA (pointwise) warps of an SM sub-partition share d when D = 2
    const size_t scr_doubles = (size_t)G * D * NT * 64;
    // shared memory: [Ke scratch (if it fits)] [rings: D x R stages x 2 pair-blocks] [queues: tasks x S x XS x 32] [mbarriers]
    // (the per-chain reduction area `red` aliases the queues, which are dead by then)
    double* kscr = a.scratch_in_smem ? smem : a.scratch + (size_t)blockIdx.x * scr_doubles;
    double* rings = smem + (a.scratch_in_smem ? scr_doubles : 0);
    double* ringA = rings + (size_t)d_rt * R * 2 * BLKP;          // A1: m~ pair-blocks (pointwise warps); A2: C~ pair-blocks
    double* ringB = ringA + (size_t)R * BLKP;                      // A1: K~ pair-blocks; A2: m~^T pair-blocks (DMMA warps)
    double* xq_base = rings + (size_t)D * R * 2 * BLKP;
    double* xq = xq_base + (size_t)task * S * XS * 32 + lane;
    unsigned long long* mbars = reinterpret_cast<unsigned long long*>(xq_base + (size_t)ntask * S * XS * 32);
    unsigned long long* afull = mbars + d_rt * 4 * R;            // ring A stages of this dimension
    unsigned long long* aempty = afull + R;
    unsigned long long* bfull = aempty + R;                      // ring B
    unsigned long long* bempty = bfull + R;
    unsigned long long* q1full = mbars + 4 * R * D + task * 4 * S;   // A1: P -> C (e tiles ready)
    unsigned long long* q1empty = q1full + S;                        // A1: C -> P (stage consumed)
    unsigned long long* q2full = q1empty + S;                        // A2: C -> P (Cx, m^T Ke tiles ready)
    unsigned long long* q2empty = q2full + S;                        // A2: P -> C (stage consumed)
    double* red = xq_base;                                           // [G*8][D][RED]
#ifdef MAGI_POISON_SMEM
    {   // debug: every shared-memory word starts as NaN, so that a read of a never-written location shows up deterministically
        const size_t tot = (a.scratch_in_smem ? scr_doubles : 0) + (size_t)D * R * 2 * BLKP + (size_t)ntask * S * XS * 32;
        for (size_t i = threadIdx.x; i < tot; i += blockDim.x) smem[i] = __longlong_as_double(0x7ff8000000000000LL | (long long)(MAGI_POISON_SMEM));
        __syncthreads();
    }
#endif
    {   // mbarrier init, one barrier per thread: [D][rfull R | rempty R] then [task][4 S]
        const int nring = 4 * R * D, nall = nring + 4 * S * ntask;
        for (int i = threadIdx.x; i < nall; i += blockDim.x)
            mbar_init(mbars + i, (i < nring && ((i / R) & 1)) ? G : 1);     // the "empty" barriers take one arrival per consumer warp
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const long long chain = (long long)blockIdx.x * (G * 8) + g * 8 + gid;
    const bool cvalid = chain < a.n_chains;
    const double* xp = a.params + (cvalid ? chain : (long long)a.n_chains - 1) * a.pitch;
    const double inv_b3 = a.inv_beta[2];
    long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0;
    long long wq = 0, wfull = 0;                 // MAGI_DBG_WAITS: cycles spent waiting on the queues / the fragment ring
    if (a.dbg) tk0 = clock64();

    auto tile_ok = [&](int J) { return J >= 0 && J < NT; };
    auto ld2 = [&](const double* base, int J, double& v0, double& v1) {   // the two values this lane owns in tile J (0 outside)
        const int t0 = 8 * J + q, t1 = t0 + 4;
        v0 = (t0 >= 0 && t0 < n) ? base[t0] : 0.0;
        v1 = (t1 >= 0 && t1 < n) ? base[t1] : 0.0;
    };
    const int NP = (NT + 1) / 2;                    // pair-blocks per (view, dimension): tiles (2p, 2p + 1)
    auto pair_ok = [&](int pp) { return pp >= 0 && pp < NP; };
    // One elected thread streams pair-block `src` (null: nothing to load, the barrier still completes) into stage r mod R of
    // a ring, after every consumer warp has released use r - R.  The producer duty rotates over the G consumer warps of the
    // dimension (use r is issued by warp g = r mod G), so that no warp is systematically slower than the ones it shares the
    // ring with.  No proxy fence: the stage was only READ through the generic proxy, before the consumers' arrivals.
    auto ring_issue = [&](int r, double* buf, unsigned long long* full, unsigned long long* empty, const double* src, double* buf2, const double* src2) {
        if ((r & (G - 1)) != g) return;                                   // warp-uniform (G is a power of two: banded_pick_config)
        if (lane == 0) {
            const int st = r % R;
            if (r >= R) mbar_wait(empty + st, ((r / R) - 1) & 1);
            const unsigned nb = (src != nullptr) + (src2 != nullptr);
            mbar_expect_tx(full + st, nb * BLKP * 8);
            if (src) tma_bulk_g2s(buf + (size_t)st * BLKP, src, BLKP * 8, full + st);
            if (src2) tma_bulk_g2s(buf2 + (size_t)st * BLKP, src2, BLKP * 8, full + st);
        }
        __syncwarp();
    };

    // =========================== DMMA warp ===========================
    auto c_warp = [&]() {                          // one code path for every dimension (d only enters addresses)
        const int d = d_rt;
        const double* xd = xp + (size_t)d * n;
        const double* ft1 = a.fragtab + ((size_t)(1 * D + d) * NP) * BLKP;   // C~
        const double* ft2 = a.fragtab + ((size_t)(2 * D + d) * NP) * BLKP;   // K~
        const double* ft3 = a.fragtab + ((size_t)(3 * D + d) * NP) * BLKP;   // m~^T
        // ring B use r: A1 K~ pair r (r < NP); A2 step u = r - NP: C~ pair into ring A's buffer and m~^T pair into ring B's,
        // both signalled on ring B's barriers (the pointwise warps no longer use ring A then)
        auto issue_b = [&](int r) {
            if (r < NP) ring_issue(r, ringB, bfull, bempty, ft2 + (size_t)r * BLKP, nullptr, nullptr);
            else if (r < 2 * NP) {
                const int pc = r - NP;
                ring_issue(r, ringB, bfull, bempty, pair_ok(pc) ? ft3 + (size_t)pc * BLKP : nullptr, ringA, pair_ok(pc) ? ft1 + (size_t)pc * BLKP : nullptr);
            }
        };
        issue_b(0);
        if (NP > 1) issue_b(1);
        double acc_eke = 0.0, acc_xcx = 0.0;
        // ---------------- A1: Ke = K~ e (likelihoods.jl:132); e tiles come from the pointwise warp ----------------
        {
            double ew[W2E];
#pragma unroll
            for (int i = 0; i < W2E; ++i) ew[i] = 0.0;
            double* ks = kscr + ((size_t)(g * D + d) * NT) * 64 + lane;
            auto a1_step = [&](int i, auto in_, auto vb_) {
                constexpr bool IN = decltype(in_)::value, VB = decltype(vb_)::value;   // IN: an e pair arrives; VB: a K~ pair is computed
                const int pb = i - LAGH, st = pb % R, qs = i % S;
                double e4[4] = {0.0, 0.0, 0.0, 0.0};
#ifdef MAGI_DBG_WAITS
                long long w0 = clock64();
#endif
                if constexpr (IN && VB) mbar_wait2(q1full + qs, (i / S) & 1, bfull + st, (pb / R) & 1);
                else if constexpr (IN) mbar_wait(q1full + qs, (i / S) & 1);
                else if constexpr (VB) mbar_wait(bfull + st, (pb / R) & 1);
#ifdef MAGI_DBG_WAITS
                wfull += clock64() - w0;
#endif
                if constexpr (IN) {
                    const double* xs = xq + (size_t)qs * XS * 32;
                    e4[0] = xs[0]; e4[1] = xs[32]; e4[2] = xs[64]; e4[3] = xs[96];
                }
#pragma unroll
                for (int j = 0; j < W2E - 4; ++j) ew[j] = ew[j + 4];
#pragma unroll
                for (int j = 0; j < 4; ++j) ew[W2E - 4 + j] = e4[j];
                if constexpr (VB) {
                    const double2* fr = reinterpret_cast<const double2*>(ringB + (size_t)st * BLKP) + lane;
                    double k[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                    // one 16-byte LDS = chunk hh of BOTH tiles of the pair: the two DMMAs it feeds belong to different accumulate chains
#pragma unroll
                    for (int hh = 0; hh < NCH; ++hh) {
                        const double2 fb = fr[hh * 32];
                        dmma884(k[0][0], k[0][1], ew[hh], fb.x);  dmma884(k[1][0], k[1][1], ew[hh + 2], fb.y);
                    }
                    __syncwarp();
                    if (lane == 0) { if (IN) mbar_arrive(q1empty + qs); mbar_arrive(bempty + st); }
                    if (pb + 2 < NP) issue_b(pb + 2);         // (the A2 uses wait for the block barrier: the pointwise warps may still read ring A)
#pragma unroll
                    for (int tt = 0; tt < 2; ++tt) {
                        if (tile_ok(2 * pb + tt)) {
                            ks[(size_t)(2 * pb + tt) * 64] = k[tt][0];
                            ks[(size_t)(2 * pb + tt) * 64 + 32] = k[tt][1];
                            acc_eke += ew[HB + 2 * tt] * k[tt][0];             // likelihoods.jl:146
                            acc_eke += ew[HB + 2 * tt + 1] * k[tt][1];
                        }
                    }
                } else if constexpr (IN) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(q1empty + qs);
                }
            };
            // e pairs arrive for i in [0, NP), K~ pairs are computed for i in [LAGH, NP + LAGH): one loop per combination
            const int iend = NP + LAGH;
            int bp[4] = {0, min(LAGH, iend), min(NP, iend), min(NP + LAGH, iend)};
            if (bp[1] > bp[2]) { const int t = bp[1]; bp[1] = bp[2]; bp[2] = t; }
            if (bp[2] > bp[3]) { const int t = bp[2]; bp[2] = bp[3]; bp[3] = t; }
            if (bp[1] > bp[2]) { const int t = bp[1]; bp[1] = bp[2]; bp[2] = t; }
#pragma unroll 1
            for (int seg = 0; seg < 4; ++seg) {
                const int i0 = bp[seg], i1 = seg < 3 ? bp[seg + 1] : iend;
                if (i0 >= i1) continue;
                const bool in = i0 < NP, vb = (i0 >= LAGH) && (i0 < NP + LAGH);
                if (in && vb) {
#pragma unroll 1
                    for (int i = i0; i < i1; ++i) a1_step(i, std::true_type{}, std::true_type{});
                } else if (in) {
#pragma unroll 1
                    for (int i = i0; i < i1; ++i) a1_step(i, std::true_type{}, std::false_type{});
                } else if (vb) {
#pragma unroll 1
                    for (int i = i0; i < i1; ++i) a1_step(i, std::false_type{}, std::true_type{});
                } else {
#pragma unroll 1
                    for (int i = i0; i < i1; ++i) a1_step(i, std::false_type{}, std::false_type{});
                }
            }
            acc_eke = quad_sum(acc_eke);
        }
        if (a.dbg) tk1 = clock64();
        __syncthreads();                                     // Ke of every dimension is in the scratch
        if (a.dbg) tk2 = clock64();
        // ---------------- A2 ----------------
        {
            double xw[W2], kw[W2], nf[4], nk[4];
#pragma unroll
            for (int i = 0; i < W2; ++i) { xw[i] = 0.0; kw[i] = 0.0; }
            const double* ks = kscr + ((size_t)(g * D + d) * NT) * 64 + lane;
            auto load_feed = [&](int u, double (&fx)[4], double (&fk)[4]) {
                ld2(xd, 2 * u - HO, fx[0], fx[1]); ld2(xd, 2 * u - HO + 1, fx[2], fx[3]);      // window heads of step u
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) {
                    const int J = 2 * u - HO + tt;
                    const bool ok = tile_ok(J);
                    fk[2 * tt] = ok ? ks[(size_t)(ok ? J : 0) * 64] : 0.0;
                    fk[2 * tt + 1] = ok ? ks[(size_t)(ok ? J : 0) * 64 + 32] : 0.0;
                }
            };
            // the windows start out filled up to the heads of step LAGH (the first step with an output pair): no lead-in steps
#pragma unroll
            for (int k = 0; k <= LAGH; ++k) {
                load_feed(LAGH - k, nf, nk);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (W2 - 4 - 4 * k + i >= 0) { xw[W2 - 4 - 4 * k + i] = nf[i]; kw[W2 - 4 - 4 * k + i] = nk[i]; }
            }
            // ping-pong partner: the other DMMA warp on this SM sub-partition (warp ^ 4), named barriers 1..8
            const bool pp_partner = ((warp ^ 4) < ntask) && a.H == 0, pp_first = warp < 4;
            const int pp_mine = 1 + 2 * (warp & 3) + (pp_first ? 0 : 1), pp_other = 1 + 2 * (warp & 3) + (pp_first ? 1 : 0);
            auto a2_step = [&](int pc) {                     // output pair pc = tiles (2 pc, 2 pc + 1); window heads of step u = pc + LAGH
                const int u = pc + LAGH, r = NP + pc;
                const int st = r % R, qs = pc % S;
                double* xs = xq + (size_t)qs * XS * 32;
#ifdef MAGI_DBG_WAITS
                long long w0 = clock64();
#endif
                // fragments of this step; the pointwise warp has read the tiles of step u - S
                if (pc >= S) mbar_wait2(bfull + st, (r / R) & 1, q2empty + qs, ((pc / S) - 1) & 1);
                else mbar_wait(bfull + st, (r / R) & 1);
#ifdef MAGI_DBG_WAITS
                wfull += clock64() - w0;
#endif
                load_feed(u + 1, nf, nk);                        // x and Ke tiles (2u+2, 2u+3), entered at the end of the step
                const double2* fra = reinterpret_cast<const double2*>(ringA + (size_t)st * BLKP) + lane;
                const double2* frb = reinterpret_cast<const double2*>(ringB + (size_t)st * BLKP) + lane;
                double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, um[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                {
                    // The two DMMA warps of an SM sub-partition take turns on the FP64 tensor pipe: while one issues its 48 DMMAs
                    // (16 clk each, alone on the pipe) the other does its non-DMMA part of the step.  Left alone they run in
                    // lock-step (same ring barrier releases both), share the pipe during the DMMA blocks and leave it idle
                    // during the rest.
                    if (pp_partner) named_barrier(pp_mine, 64);
#pragma unroll
                    for (int hh = 0; hh < NCH; ++hh) {
                        const double2 fa = fra[hh * 32], fb = frb[hh * 32];
                        dmma884(c[0][0], c[0][1], xw[hh], fa.x);    dmma884(c[1][0], c[1][1], xw[hh + 2], fa.y);     // likelihoods.jl:133
                        dmma884(um[0][0], um[0][1], kw[hh], fb.x);  dmma884(um[1][0], um[1][1], kw[hh + 2], fb.y);   // likelihoods.jl:192
                    }
                    if (pp_partner && (pp_first || pc + 1 < NP)) named_arrive(pp_other, 64);
                }
                // the DMMA warp has slack in this phase: it combines the two products (likelihoods.jl:186,194) and accumulates
                // x.Cx (:150; x and Cx are both zero outside the time axis), the pointwise warp gets one value per point
#pragma unroll
                for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                    for (int pt = 0; pt < 2; ++pt) {
                        acc_xcx += xw[HB + 2 * tt + pt] * c[tt][pt];
                        xs[(2 * tt + pt) * 32] = um[tt][pt] - c[tt][pt];      // 1/beta1, 1/beta2 are folded into the K~ and C~ fragment tables
                    }
                __syncwarp();
                if (lane == 0) { mbar_arrive(q2full + qs); mbar_arrive(bempty + st); }
                issue_b(r + 2);
#pragma unroll
                for (int i = 0; i < W2 - 4; ++i) { xw[i] = xw[i + 4]; kw[i] = kw[i + 4]; }
#pragma unroll
                for (int i = 0; i < 4; ++i) { xw[W2 - 4 + i] = nf[i]; kw[W2 - 4 + i] = nk[i]; }
            };
            issue_b(NP); issue_b(NP + 1);                      // (after the block barrier: ring A's buffer is free now)
            if (pp_partner && !pp_first && NP > 0) named_arrive(pp_other, 64);      // the first warp of the pair may start
#pragma unroll 1
            for (int pc = 0; pc < NP; ++pc) a2_step(pc);
        }
        if (a.dbg) tk3 = clock64();
        __syncthreads();                                     // the queues are dead: `red` may overwrite them
        acc_xcx = quad_sum(acc_xcx);
        if (q == 0) { red[((size_t)(g * 8 + gid) * D + d) * RED + 0] = acc_eke; red[((size_t)(g * 8 + gid) * D + d) * RED + 1] = acc_xcx; }
    };

    // =========================== pointwise warp ===========================
    auto p_warp = [&](auto dconst) {
        constexpr int d = decltype(dconst)::value;
        double th[M::KX];
#pragma unroll
        for (int i = 0; i < K; ++i) th[i] = xp[(size_t)n * D + i];
        M::prepare(th);
        // ---------------- A1: mx = m~ x_d (likelihoods.jl:129), e = f(x, theta) - mx (:130) -> queue, ahead of the DMMA warp ----------------
        // step pa: output tiles (Ja, Ja+1) = m~ pair pa (Ja = 2 pa); the x window head is at tiles (2u - HO, 2u - HO + 1), u = pa + LAGH
        {
            const double* xd = xp + (size_t)d * n;
            const double* ft0 = a.fragtab + ((size_t)(0 * D + d) * NP) * BLKP;   // m~
            auto issue_a = [&](int r) { if (r < NP) ring_issue(r, ringA, afull, aempty, ft0 + (size_t)r * BLKP, nullptr, nullptr); };
            issue_a(0); issue_a(1);
            double xw[W2], nf[4];
#pragma unroll
            for (int i = 0; i < W2; ++i) xw[i] = 0.0;
            // the window starts out filled up to the heads of step LAGH (the first step with an output pair): no lead-in steps
#pragma unroll
            for (int k = 0; k <= LAGH; ++k) {
                double h4[4];
                ld2(xd, 2 * (LAGH - k) - HO, h4[0], h4[1]); ld2(xd, 2 * (LAGH - k) - HO + 1, h4[2], h4[3]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (W2 - 4 - 4 * k + i >= 0) xw[W2 - 4 - 4 * k + i] = h4[i];
            }
            auto a1_step = [&](int pa) {
                constexpr bool VA = true;
                const int u = pa + LAGH, Ja = 2 * pa, st = pa % R, qs = pa % S;
                if constexpr (VA) {
#ifdef MAGI_DBG_WAITS
                    long long w0 = clock64();
#endif
                    if (pa >= S) mbar_wait2(afull + st, (pa / R) & 1, q1empty + qs, ((pa / S) - 1) & 1);   // fragments; queue stage free
                    else mbar_wait(afull + st, (pa / R) & 1);
#ifdef MAGI_DBG_WAITS
                    wq += clock64() - w0;
#endif
                }
                ld2(xd, 2 * u + 2 - HO, nf[0], nf[1]); ld2(xd, 2 * u + 3 - HO, nf[2], nf[3]);    // window feed of step u+1 (entered at the end)
                double xo[2][2][D];                  // the other components at the output points (this one is in the window)
#pragma unroll
                for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                    for (int dd = 0; dd < D; ++dd)
                        if (dd != d) ld2(xp + (size_t)dd * n, (VA && tile_ok(Ja + tt)) ? Ja + tt : -4, xo[tt][0][dd], xo[tt][1][dd]);
                if constexpr (VA) {
                    const double2* fr = reinterpret_cast<const double2*>(ringA + (size_t)st * BLKP) + lane;
                    double m[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
                    for (int hh = 0; hh < NCH; ++hh) {
                        const double2 fa = fr[hh * 32];
                        dmma884(m[0][0], m[0][1], xw[hh], fa.x);  dmma884(m[1][0], m[1][1], xw[hh + 2], fa.y);
                    }
                    double ev[2][2];
#pragma unroll
                    for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                        for (int pt = 0; pt < 2; ++pt) {
                            const int t = 8 * (Ja + tt) + q + 4 * pt;
                            xo[tt][pt][d] = xw[HB + 2 * tt + pt];
                            const double v = M::f(d, xo[tt][pt], th) - m[tt][pt];
                            ev[tt][pt] = (tile_ok(Ja + tt) && t < n) ? v : 0.0;
                        }
                    double* xs = xq + (size_t)qs * XS * 32;
                    xs[0] = ev[0][0]; xs[32] = ev[0][1]; xs[64] = ev[1][0]; xs[96] = ev[1][1];
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(q1full + qs); mbar_arrive(aempty + st); }
                    issue_a(pa + 2);
                }
#pragma unroll
                for (int j = 0; j < W2 - 4; ++j) xw[j] = xw[j + 4];
#pragma unroll
                for (int j = 0; j < 4; ++j) xw[W2 - 4 + j] = nf[j];
            };
#pragma unroll 1
            for (int pa = 0; pa < NP; ++pa) a1_step(pa);
        }
        // A2 prologue that does not depend on the Ke scratch
        double acc_sse = 0.0;
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
        const double obs_scale = (1.0 / (sigma_d * sigma_d)) * inv_b3;
        const double* yd = a.yobs + (size_t)d * n;
        double* gout = (a.grad != nullptr && cvalid) ? a.grad + chain * a.pitch + (size_t)d * n : nullptr;
        double nxa[2][2][D], nyv[2][2];
        // INTERIOR steps (both output tiles inside the time axis, all 16 times < n) run without bounds predicates and selects
        auto load_a2_int = [&](int u, double (&X)[2][2][D], double (&Y)[2][2]) {
            const int Jc = 2 * u;
            const double* px = xp + 8 * Jc + q;
            const double* py = yd + 8 * Jc + q;
#pragma unroll
            for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                for (int pt = 0; pt < 2; ++pt) {
#pragma unroll
                    for (int dd = 0; dd < D; ++dd) X[tt][pt][dd] = px[(size_t)dd * n + 8 * tt + 4 * pt];
                    Y[tt][pt] = py[8 * tt + 4 * pt];
                }
        };
        auto load_a2 = [&](int u, double (&X)[2][2][D], double (&Y)[2][2]) {
            const int Jc = 2 * u;
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                const int J = tile_ok(Jc + tt) ? Jc + tt : -4;
#pragma unroll
                for (int dd = 0; dd < D; ++dd) ld2(xp + (size_t)dd * n, J, X[tt][0][dd], X[tt][1][dd]);
                ld2(yd, J, Y[tt][0], Y[tt][1]);
            }
        };
        auto interior = [&](int u) { const int Jc = 2 * u; return u < NP && Jc >= 0 && 8 * (Jc + 1) + 7 < n; };
        load_a2(0, nxa, nyv);
        __syncthreads();
        // ---------------- A2: pointwise gradient; step u handles the output pair u = tiles (2u, 2u + 1) ----------------
        auto a2_step = [&](int u, auto int_) {
            constexpr bool INTERIOR = decltype(int_)::value;     // this step AND the next one are interior
            const int Jc = 2 * u, qs = u % S;
            double xa[2][2][D], wv[2][2][D], yv[2][2];
#pragma unroll
            for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                for (int pt = 0; pt < 2; ++pt) {
#pragma unroll
                    for (int dd = 0; dd < D; ++dd) xa[tt][pt][dd] = nxa[tt][pt][dd];
                    yv[tt][pt] = nyv[tt][pt];
                }
            if constexpr (INTERIOR) load_a2_int(u + 1, nxa, nyv);
            else if (u + 1 < NP) load_a2(u + 1, nxa, nyv);
            if constexpr (INTERIOR) {
                const double* wsrc = kscr + ((size_t)(g * D) * NT + Jc) * 64 + lane;
#pragma unroll
                for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                    for (int dd = 0; dd < D; ++dd) {
                        wv[tt][0][dd] = wsrc[((size_t)dd * NT + tt) * 64];                  // likelihoods.jl:201 (the scratch holds Ke / beta1)
                        wv[tt][1][dd] = wsrc[((size_t)dd * NT + tt) * 64 + 32];
                    }
            } else {
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) {
                    const int J = tile_ok(Jc + tt) ? Jc + tt : -4;
#pragma unroll
                    for (int dd = 0; dd < D; ++dd) {
                        const double* wsrc = kscr + ((size_t)(g * D + dd) * NT + (J < 0 ? 0 : J)) * 64 + lane;
                        wv[tt][0][dd] = (J < 0) ? 0.0 : wsrc[0];           // likelihoods.jl:201
                        wv[tt][1][dd] = (J < 0) ? 0.0 : wsrc[32];
                    }
                }
            }
            double* xs = xq + (size_t)qs * XS * 32;
#ifdef MAGI_DBG_WAITS
            long long w0 = clock64();
#endif
            mbar_wait(q2full + qs, (u / S) & 1);
#ifdef MAGI_DBG_WAITS
            wq += clock64() - w0;
#endif
            const double gb[2][2] = {{xs[0], xs[32]}, {xs[64], xs[96]}};      // m~^T Ke / beta1 - C~ x / beta2 at this lane's four points
            __syncwarp();
            if (lane == 0) mbar_arrive(q2empty + qs);            // the stage is free again while the gradient is computed
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
#pragma unroll
                for (int pt = 0; pt < 2; ++pt) {
                    const int t = 8 * (Jc + tt) + q + 4 * pt;
                    const bool valid = INTERIOR || (tile_ok(Jc + tt) && t < n);
                    const double* xv = xa[tt][pt];
                    const double* w = wv[tt][pt];
                    const double xdv = xv[d], wd = valid ? w[d] : 0.0;
                    const double y = yv[tt][pt];
                    const bool fin = valid && isfinite(y);         // likelihoods.jl:123
                    const double e0 = fin ? xdv - y : 0.0;
                    double gv = gb[tt][pt] - e0 * obs_scale;       // likelihoods.jl:179 (e0 = 0 when the observation is missing), :186, :194
                    M::jx_col_sub(d, xv, th, w, gv);               // likelihoods.jl:214-216
                    M::jth_row_sub(d, xv, th, wd, gth);            // likelihoods.jl:219-221
                    acc_sse += e0 * e0;                            // likelihoods.jl:139,234
                    bad |= valid && !isfinite(gv);
                    if (valid && gout != nullptr) gout[t] = gv;
                }
            }
        };
        {
            int u = 0;
#pragma unroll 1
            for (; u < NP && !(interior(u) && interior(u + 1)); ++u) a2_step(u, std::false_type{});
#pragma unroll 1
            for (; u < NP && interior(u) && interior(u + 1); ++u) a2_step(u, std::true_type{});
#pragma unroll 1
            for (; u < NP; ++u) a2_step(u, std::false_type{});
        }
        acc_sse = quad_sum(acc_sse);
#pragma unroll
        for (int i = 0; i < K; ++i) gth[i] = quad_sum(gth[i]);
        const unsigned badm = __ballot_sync(0xffffffffu, bad);
        __syncthreads();                                     // the queues are dead: `red` may overwrite them
        if (q == 0) {
            double* r = red + ((size_t)(g * 8 + gid) * D + d) * RED;
            r[2] = acc_sse;
            r[3] = (((badm >> (gid * 4)) & 0xfu) && a.grad != nullptr) ? 1.0 : 0.0;    // value-only calls look at ll alone (interface.jl:155-160)
#pragma unroll
            for (int i = 0; i < K; ++i) r[4 + i] = gth[i];
        }
    };

    if (role == 0) c_warp();
    else dispatch_dim<D>(d_rt, p_warp);
    __syncthreads();
    if (a.dbg && lane == 0) {
        long long* o = a.dbg + ((size_t)blockIdx.x * ntask + task) * 8;
        if (role == 0) { o[0] = tk1 - tk0; o[1] = tk2 - tk1; o[2] = tk3 - tk2; o[3] = clock64() - tk3; o[4] = wfull; o[5] = wq; }
        else o[7] = wq;
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
                {
                    const double* r = red + ((size_t)threadIdx.x * D + d) * RED;
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
                ll += -0.5 * eke;                                     // :146-147 (1/beta1 folded into K~)
                ll += -0.5 * xcx;                                     // :150-151 (1/beta2 folded into C~)
                gsig[d] = (s > 0 && nobs > 0) ? (sse / s2 - nobs) / (s * a.beta[2]) : 0.0;   // :229-246
                if (gp) bad |= !isfinite(gsig[d]);
            }
#pragma unroll
            for (int i = 0; i < K; ++i) if (gp) bad |= !isfinite(gthf[i]);
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

template <int MODEL, int HB>
static cudaError_t launch_one(const BandedArgs& a, int DW, size_t smem_bytes, cudaStream_t st) {
    auto kern = banded_logpost_kernel<MODEL, HB>;
    static PerDeviceOnce once;       // per instantiation
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
    }
    const int threads = 2 * a.G * DW * 32;      // one DMMA warp and one pointwise warp per (chain-group, dimension) task
    const int blocks = (a.n_chains + a.G * 8 - 1) / (a.G * 8);
    kern<<<blocks, threads, smem_bytes, st>>>(a);
    return cudaGetLastError();
}

template <int MODEL>
cudaError_t launch_model(const BandedArgs& a, int HB, int DW, size_t smem_bytes, cudaStream_t st) {
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


}  // namespace magi
