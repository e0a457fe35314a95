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
//   phase A1:  mx = m~ x_d;  e = f_d(x, theta) - mx;  Ke = K~ e  -> Ke scratch; sum e.Ke
//   phase A2:  Cx = C~ x_d;  mt = m~^T Ke_d;  pointwise gradient incl. the ODE Jacobian terms, which need Ke of ALL
//              dimensions at the same time point (hence the block-wide barrier in between)
//   final (per chain):  log-likelihood assembly in the reference's term order, sigma gradient, log-sigma transform and the
//              per-chain -Inf / zero-gradient guards (interface.jl:179-264).
// Warp roles, exchange protocol and the measured behaviour: see the comment above the kernel and DESIGN.md section 4.
#pragma once
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include "magi_common.cuh"
#include "ode_models.cuh"

namespace magi {

// ------------------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double quad_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// ---- TMA bulk copy + mbarrier helpers (fragment ring) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned it = 0; it < (1u << 26); ++it) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void named_barrier(int id, int nthreads) { asm volatile("bar.sync %0, %1;\n" :: "r"(id), "r"(nthreads) : "memory"); }

// One block = G chain-groups (8 chains each) x DW dimension slots, one warp per (group, dim) task, sweeping the time axis
// TWO 8-time tiles per step (halves the per-step overhead -- window shifts, ring bookkeeping, barriers, address math -- and
// gives the scheduler four independent DMMA accumulate chains plus two independent pointwise evaluations per step).
//   A1: mx = m~ x_d, e = f - mx (sliding register windows), Ke = K~ e -> Ke scratch (shared memory); sum e.Ke
//   A2: Cx = C~ x_d ; mt = m~^T Ke_d ; pointwise gradient incl. the ODE Jacobian terms (need Ke of all dimensions)
// The four table-fragment blocks a step needs are staged ONCE per dimension in a double-buffered cp.async ring in shared
// memory, shared by the G warps of that dimension (one named barrier per step), and read with 16-byte LDS (two chunks per
// load): each fragment block leaves L2 once per block.
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}

// calls f(std::integral_constant<int, d>) for the runtime (warp-uniform) dimension d: the model functors then see a
// compile-time dimension and compile to straight-line code
template <int D, class F> __device__ __forceinline__ void dispatch_dim(int d, F& f) {
    if constexpr (D >= 1) { if (d == 0) { f(std::integral_constant<int, 0>{}); return; } }
    if constexpr (D >= 2) { if (d == 1) { f(std::integral_constant<int, 1>{}); return; } }
    if constexpr (D >= 3) { if (d == 2) { f(std::integral_constant<int, 2>{}); return; } }
    if constexpr (D >= 4) { if (d == 3) { f(std::integral_constant<int, 3>{}); return; } }
    if constexpr (D >= 5) { if (d == 4) { f(std::integral_constant<int, 4>{}); return; } }
}

// Warp-specialised K1.  One block = G chain-groups (8 chains each) x D dimensions = G*D tasks; every task has a DMMA warp
// ("C") and a pointwise warp ("P"), 2*G*D <= 16 warps, <= 128 registers.  The C warp owns the operand windows (x, e / Ke in
// registers, sliding two 8-time tiles per step) and issues nothing but fragment LDS, window moves and DMMAs; the P warp does
// every global load and all scalar FP64 work (ODE right-hand side, Jacobian terms, reductions, gradient stores) and hands
// tiles back through a double-buffered exchange area in shared memory, with one mbarrier per direction and stage:
//   A1  C(u): mx tiles (Ja, Ja+1) -> P        P(u): e = f(x, theta) - mx, x feed for step u+2 -> C(u+2)
//       C(u): Ke tiles (Jb, Jb+1) = K~ e from the e window (lags two steps behind mx) -> Ke scratch; sum e.Ke
//   A2  C(u): Cx, m^T Ke tiles (Jc, Jc+1) -> P    P(u): gradient incl. Jacobian terms, stores, reductions; x feed -> C(u+2)
// The scalar FP64 work shares the FP64 unit with the DMMAs, but it no longer sits on the DMMA warps' in-order critical path
// (measured before the split: a step's non-DMMA section took ~1500 cycles against 768 cycles of DMMA issue).
// Fragment blocks arrive by TMA bulk copy into a double-buffered ring per dimension, shared by the G DMMA warps of that
// dimension (named barrier per step), one 16-byte LDS per two chunks.
template <int MODEL, int HB>
__global__ void __launch_bounds__(512, 1) banded_logpost_kernel(const BandedArgs a) {
    using M = Ode<MODEL>;
    constexpr int D = M::D, K = M::K;
    constexpr int NCH = 2 * HB + 2, LAGT = (HB + 1) / 2, WN = 2 * LAGT + 2 + HB, W2 = WN + 2;
    constexpr int RED = 4 + K;   // e.Ke, x.Cx, sse, bad flag, theta-gradient partials
    constexpr int BLK = NCH * 32;               // doubles per fragment block (one view, one tile)
    constexpr int XS = 12;                      // exchange slots per stage: A1 C->P 0..3 (mx), P->C 4..7 (e), 8..11 (x feed);
                                                //                           A2 C->P 0..3 (Cx), 4..7 (m^T Ke), P->C 8..11 (x feed)
    extern __shared__ __align__(128) double smem[];
    const int NT = a.NT, n = a.n, G = a.G;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
    const int ntask = G * D;
    const int role = warp / ntask;               // 0: DMMA warp (C), 1: pointwise warp (P)
    const int task = warp % ntask, g = task % G, d_rt = task / G;
    const size_t scr_doubles = (size_t)G * D * NT * 64;
    // shared memory: [Ke scratch (if it fits)] [rings: D x 2 stages x 4 blocks] [exchange: tasks x 2 x XS x 32] [mbarriers] [red]
    double* kscr = a.scratch_in_smem ? smem : a.scratch + (size_t)blockIdx.x * scr_doubles;
    double* rings = smem + (a.scratch_in_smem ? scr_doubles : 0);
    double* ring = rings + (size_t)d_rt * 2 * 4 * BLK;
    double* xch = rings + (size_t)D * 2 * 4 * BLK + (size_t)task * 2 * XS * 32 + lane;
    unsigned long long* mbars = reinterpret_cast<unsigned long long*>(rings + (size_t)D * 2 * 4 * BLK + (size_t)ntask * 2 * XS * 32);
    unsigned long long* full = mbars + d_rt * 2;                 // ring stages of this dimension
    unsigned long long* c2p = mbars + 2 * D + task * 4;          // [2] C -> P (tiles ready)
    unsigned long long* p2c = c2p + 2;                           // [2] P -> C (results ready / stage consumed)
    double* red = reinterpret_cast<double*>(mbars + 2 * D + 4 * ntask);   // [G*8][D][RED]
    const int ring_threads = G * 32;
    const bool ring_leader = (role == 0 && g == 0 && lane == 0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * D + 4 * ntask; ++i) mbar_init(mbars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    unsigned ring_use = 0;                       // ring stages consumed so far (parity tracking), C warps
    unsigned xuse = 0;                           // exchange stages used so far, identical in the C and P warp of a task

    const long long chain = (long long)blockIdx.x * (G * 8) + g * 8 + gid;
    const bool cvalid = chain < a.n_chains;
    const double* xp = a.params + (cvalid ? chain : (long long)a.n_chains - 1) * a.pitch;
    double th[M::KX];
#pragma unroll
    for (int i = 0; i < K; ++i) th[i] = xp[(size_t)n * D + i];
    M::prepare(th);
    const double inv_b1 = a.inv_beta[0], inv_b2 = a.inv_beta[1], inv_b3 = a.inv_beta[2];
    long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0;
    long long wbar = 0, wp2c = 0, wfull = 0, wc2p = 0;     // MAGI_DBG_WAITS: cycles spent in each kind of wait (A2 only)
    if (a.dbg) tk0 = clock64();

    auto stage_blocks = [&](unsigned use, const double* s0, const double* s1, const double* s2, const double* s3) {
        if (ring_leader) {
            double* dst = ring + (size_t)(use & 1) * 4 * BLK;
            const unsigned nb = (s0 != nullptr) + (s1 != nullptr) + (s2 != nullptr) + (s3 != nullptr);
            fence_proxy_async();
            mbar_expect_tx(full + (use & 1), nb * BLK * 8);
            if (s0) tma_bulk_g2s(dst, s0, BLK * 8, full + (use & 1));
            if (s1) tma_bulk_g2s(dst + BLK, s1, BLK * 8, full + (use & 1));
            if (s2) tma_bulk_g2s(dst + 2 * BLK, s2, BLK * 8, full + (use & 1));
            if (s3) tma_bulk_g2s(dst + 3 * BLK, s3, BLK * 8, full + (use & 1));
        }
    };
    auto tile_ok = [&](int J) { return J >= 0 && J < NT; };
    auto ld2 = [&](const double* base, int J, double& v0, double& v1) {   // the two values this lane owns in tile J (0 outside)
        const int t0 = 8 * J + q, t1 = t0 + 4;
        v0 = (t0 >= 0 && t0 < n) ? base[t0] : 0.0;
        v1 = (t1 >= 0 && t1 < n) ? base[t1] : 0.0;
    };
    const int uA1_end = (NT + 2 * LAGT + 4) / 2;     // last A1 step: Ke pair (2u - 4 - 2 LAGT, +1) reaches tile NT - 1
    const int uA2_end = (NT - 1 + LAGT) / 2;

    // =========================== DMMA warp ===========================
    auto c_warp = [&](auto dconst) {
        constexpr int d = decltype(dconst)::value;
        const double* xd = xp + (size_t)d * n;
        // ---------------- A1 ----------------
        {
            double xw[W2], ew[W2];
#pragma unroll
            for (int i = 0; i < W2; ++i) { xw[i] = 0.0; ew[i] = 0.0; }
            double acc_eke = 0.0;
            double* ks = kscr + ((size_t)(g * D + d) * NT) * 64 + lane;
            const double* ft0 = a.fragtab + ((size_t)(0 * D + d) * NT) * BLK;
            const double* ft2 = a.fragtab + ((size_t)(2 * D + d) * NT) * BLK;
            auto stage_step = [&](int u, unsigned use) {
                if (u <= uA1_end) {
                    const int Ja = 2 * u - LAGT, Jb = 2 * u - 4 - 2 * LAGT;
                    stage_blocks(use, tile_ok(Ja) ? ft0 + (size_t)Ja * BLK : nullptr, tile_ok(Ja + 1) ? ft0 + (size_t)(Ja + 1) * BLK : nullptr,
                                 tile_ok(Jb) ? ft2 + (size_t)Jb * BLK : nullptr, tile_ok(Jb + 1) ? ft2 + (size_t)(Jb + 1) * BLK : nullptr);
                }
            };
            named_barrier(1 + d, ring_threads);
            stage_step(0, ring_use);
            for (int u = 0; u <= uA1_end; ++u, ++ring_use, ++xuse) {
#ifdef MAGI_DBG_WAITS1
                long long w0 = clock64();
#endif
                named_barrier(1 + d, ring_threads);          // all DMMA warps of this dimension finished step u-1
#ifdef MAGI_DBG_WAITS1
                long long w1 = clock64(); wbar += w1 - w0;
#endif
                stage_step(u + 1, ring_use + 1);
                double* xs = xch + (size_t)(xuse & 1) * XS * 32;
                double fin[8];                               // e tiles of step u-2 (4) and x feed of this step (4)
                if (xuse >= 2) mbar_wait(p2c + (xuse & 1), ((xuse - 2) >> 1) & 1);
#ifdef MAGI_DBG_WAITS1
                long long w2 = clock64(); wp2c += w2 - w1;
#endif
                if (u >= 2) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) fin[i] = xs[(4 + i) * 32];
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) fin[i] = 0.0;
                    ld2(xd, 2 * u, fin[4], fin[5]); ld2(xd, 2 * u + 1, fin[6], fin[7]);
                }
#pragma unroll
                for (int i = 0; i < W2 - 4; ++i) { xw[i] = xw[i + 4]; ew[i] = ew[i + 4]; }
#pragma unroll
                for (int i = 0; i < 4; ++i) { ew[W2 - 4 + i] = fin[i]; xw[W2 - 4 + i] = fin[4 + i]; }
#ifdef MAGI_DBG_WAITS1
                long long w3 = clock64();
#endif
                mbar_wait(full + (ring_use & 1), (ring_use >> 1) & 1);
#ifdef MAGI_DBG_WAITS1
                wfull += clock64() - w3;
#endif
                const double2* fr = reinterpret_cast<const double2*>(ring + (size_t)(ring_use & 1) * 4 * BLK) + lane;
                const int Ja = 2 * u - LAGT, Jb = 2 * u - 4 - 2 * LAGT;
                const bool va = tile_ok(Ja) || tile_ok(Ja + 1), vb = tile_ok(Jb) || tile_ok(Jb + 1);
                double m[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, k[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                if (va && vb) {
#pragma unroll
                    for (int hp = 0; hp < NCH / 2; ++hp) {
                        const double2 fa0 = fr[hp * 32], fa1 = fr[BLK / 2 + hp * 32], fb0 = fr[BLK + hp * 32], fb1 = fr[3 * BLK / 2 + hp * 32];
                        dmma884(m[0][0], m[0][1], xw[2 * hp], fa0.x);         dmma884(m[1][0], m[1][1], xw[2 * hp + 2], fa1.x);
                        dmma884(k[0][0], k[0][1], ew[2 * hp], fb0.x);         dmma884(k[1][0], k[1][1], ew[2 * hp + 2], fb1.x);
                        dmma884(m[0][0], m[0][1], xw[2 * hp + 1], fa0.y);     dmma884(m[1][0], m[1][1], xw[2 * hp + 3], fa1.y);
                        dmma884(k[0][0], k[0][1], ew[2 * hp + 1], fb0.y);     dmma884(k[1][0], k[1][1], ew[2 * hp + 3], fb1.y);
                    }
                } else if (va) {
#pragma unroll
                    for (int hp = 0; hp < NCH / 2; ++hp) {
                        const double2 fa0 = fr[hp * 32], fa1 = fr[BLK / 2 + hp * 32];
                        dmma884(m[0][0], m[0][1], xw[2 * hp], fa0.x);         dmma884(m[1][0], m[1][1], xw[2 * hp + 2], fa1.x);
                        dmma884(m[0][0], m[0][1], xw[2 * hp + 1], fa0.y);     dmma884(m[1][0], m[1][1], xw[2 * hp + 3], fa1.y);
                    }
                } else if (vb) {
#pragma unroll
                    for (int hp = 0; hp < NCH / 2; ++hp) {
                        const double2 fb0 = fr[BLK + hp * 32], fb1 = fr[3 * BLK / 2 + hp * 32];
                        dmma884(k[0][0], k[0][1], ew[2 * hp], fb0.x);         dmma884(k[1][0], k[1][1], ew[2 * hp + 2], fb1.x);
                        dmma884(k[0][0], k[0][1], ew[2 * hp + 1], fb0.y);     dmma884(k[1][0], k[1][1], ew[2 * hp + 3], fb1.y);
                    }
                }
                // hand mx to the pointwise warp (likelihoods.jl:129)
                xs[0] = m[0][0]; xs[32] = m[0][1]; xs[64] = m[1][0]; xs[96] = m[1][1];
                __syncwarp();
                if (lane == 0) mbar_arrive(c2p + (xuse & 1));
                if (vb) {
#pragma unroll
                    for (int tt = 0; tt < 2; ++tt) {
                        if (tile_ok(Jb + tt)) {
                            ks[(size_t)(Jb + tt) * 64] = k[tt][0];             // likelihoods.jl:132
                            ks[(size_t)(Jb + tt) * 64 + 32] = k[tt][1];
                            acc_eke += ew[HB + 2 * tt] * k[tt][0];             // likelihoods.jl:146
                            acc_eke += ew[HB + 2 * tt + 1] * k[tt][1];
                        }
                    }
                }
            }
            acc_eke = quad_sum(acc_eke);
            if (q == 0) red[((size_t)(g * 8 + gid) * D + d) * RED + 0] = acc_eke;
        }
        if (a.dbg) tk1 = clock64();
        __syncthreads();                                     // Ke of every dimension is in the scratch
        if (a.dbg) tk2 = clock64();
        // ---------------- A2 ----------------
        {
            double xw[W2], kw[W2];
#pragma unroll
            for (int i = 0; i < W2; ++i) { xw[i] = 0.0; kw[i] = 0.0; }
            const double* ks = kscr + ((size_t)(g * D + d) * NT) * 64 + lane;
            const double* ft1 = a.fragtab + ((size_t)(1 * D + d) * NT) * BLK;
            const double* ft3 = a.fragtab + ((size_t)(3 * D + d) * NT) * BLK;
            auto stage_step = [&](int u, unsigned use) {
                if (u <= uA2_end) {
                    const int Jc = 2 * u - LAGT;
                    stage_blocks(use, tile_ok(Jc) ? ft1 + (size_t)Jc * BLK : nullptr, tile_ok(Jc + 1) ? ft1 + (size_t)(Jc + 1) * BLK : nullptr,
                                 tile_ok(Jc) ? ft3 + (size_t)Jc * BLK : nullptr, tile_ok(Jc + 1) ? ft3 + (size_t)(Jc + 1) * BLK : nullptr);
                }
            };
            named_barrier(1 + d, ring_threads);
            stage_step(0, ring_use);
            for (int u = 0; u <= uA2_end; ++u, ++ring_use, ++xuse) {
#ifdef MAGI_DBG_WAITS
                long long w0 = clock64();
#endif
                named_barrier(1 + d, ring_threads);
#ifdef MAGI_DBG_WAITS
                long long w1 = clock64(); wbar += w1 - w0;
#endif
                stage_step(u + 1, ring_use + 1);
                double* xs = xch + (size_t)(xuse & 1) * XS * 32;
                double feed[4];
                if (xuse >= 2) mbar_wait(p2c + (xuse & 1), ((xuse - 2) >> 1) & 1);     // P consumed this stage two steps ago
#ifdef MAGI_DBG_WAITS
                long long w2 = clock64(); wp2c += w2 - w1;
#endif
                if (u >= 2) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) feed[i] = xs[(8 + i) * 32];
                } else { ld2(xd, 2 * u, feed[0], feed[1]); ld2(xd, 2 * u + 1, feed[2], feed[3]); }
#pragma unroll
                for (int i = 0; i < W2 - 4; ++i) { xw[i] = xw[i + 4]; kw[i] = kw[i + 4]; }
#pragma unroll
                for (int i = 0; i < 4; ++i) xw[W2 - 4 + i] = feed[i];
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) {
                    const bool ok = tile_ok(2 * u + tt);
                    kw[W2 - 4 + 2 * tt] = ok ? ks[(size_t)(2 * u + tt) * 64] : 0.0;
                    kw[W2 - 3 + 2 * tt] = ok ? ks[(size_t)(2 * u + tt) * 64 + 32] : 0.0;
                }
#ifdef MAGI_DBG_WAITS
                long long w3 = clock64();
#endif
                mbar_wait(full + (ring_use & 1), (ring_use >> 1) & 1);
#ifdef MAGI_DBG_WAITS
                wfull += clock64() - w3;
#endif
                const double2* fr = reinterpret_cast<const double2*>(ring + (size_t)(ring_use & 1) * 4 * BLK) + lane;
                const int Jc = 2 * u - LAGT;
                double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, um[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                if (tile_ok(Jc) || tile_ok(Jc + 1)) {
#pragma unroll
                    for (int hp = 0; hp < NCH / 2; ++hp) {
                        const double2 fa0 = fr[hp * 32], fa1 = fr[BLK / 2 + hp * 32], fb0 = fr[BLK + hp * 32], fb1 = fr[3 * BLK / 2 + hp * 32];
                        dmma884(c[0][0], c[0][1], xw[2 * hp], fa0.x);          dmma884(c[1][0], c[1][1], xw[2 * hp + 2], fa1.x);      // likelihoods.jl:133
                        dmma884(um[0][0], um[0][1], kw[2 * hp], fb0.x);        dmma884(um[1][0], um[1][1], kw[2 * hp + 2], fb1.x);    // likelihoods.jl:192
                        dmma884(c[0][0], c[0][1], xw[2 * hp + 1], fa0.y);      dmma884(c[1][0], c[1][1], xw[2 * hp + 3], fa1.y);
                        dmma884(um[0][0], um[0][1], kw[2 * hp + 1], fb0.y);    dmma884(um[1][0], um[1][1], kw[2 * hp + 3], fb1.y);
                    }
                }
                xs[0] = c[0][0]; xs[32] = c[0][1]; xs[64] = c[1][0]; xs[96] = c[1][1];
                xs[128] = um[0][0]; xs[160] = um[0][1]; xs[192] = um[1][0]; xs[224] = um[1][1];
                __syncwarp();
                if (lane == 0) mbar_arrive(c2p + (xuse & 1));
            }
        }
        if (a.dbg) tk3 = clock64();
    };

    // =========================== pointwise warp ===========================
    auto p_warp = [&](auto dconst) {
        constexpr int d = decltype(dconst)::value;
        const double* xd = xp + (size_t)d * n;
        // ---------------- A1: e = f(x, theta) - mx ----------------
        // (global loads of step u+1 are issued before step u is processed: the load latency is off the exchange's critical path)
        {
            double xa[2][2][D], feed[4], nxa[2][2][D], nfeed[4];
            auto load_a1 = [&](int u, double (&X)[2][2][D], double (&F)[4]) {
                const int Ja = 2 * u - LAGT;
#pragma unroll
                for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                    for (int dd = 0; dd < D; ++dd) ld2(xp + (size_t)dd * n, tile_ok(Ja + tt) ? Ja + tt : -4, X[tt][0][dd], X[tt][1][dd]);
                ld2(xd, 2 * u + 4, F[0], F[1]); ld2(xd, 2 * u + 5, F[2], F[3]);      // window feed of step u+2
            };
            load_a1(0, nxa, nfeed);
            for (int u = 0; u <= uA1_end; ++u, ++xuse) {
                const int Ja = 2 * u - LAGT;
#pragma unroll
                for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                    for (int pt = 0; pt < 2; ++pt)
#pragma unroll
                        for (int dd = 0; dd < D; ++dd) xa[tt][pt][dd] = nxa[tt][pt][dd];
#pragma unroll
                for (int i = 0; i < 4; ++i) feed[i] = nfeed[i];
                if (u < uA1_end) load_a1(u + 1, nxa, nfeed);
                double* xs = xch + (size_t)(xuse & 1) * XS * 32;
#ifdef MAGI_DBG_WAITS1
                long long w0 = clock64();
#endif
                mbar_wait(c2p + (xuse & 1), (xuse >> 1) & 1);
#ifdef MAGI_DBG_WAITS1
                wc2p += clock64() - w0;
#endif
                const double mm[2][2] = {{xs[0], xs[32]}, {xs[64], xs[96]}};
#pragma unroll
                for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                    for (int pt = 0; pt < 2; ++pt) {
                        const int t = 8 * (Ja + tt) + q + 4 * pt;
                        const double ev = M::f(d, xa[tt][pt], th) - mm[tt][pt];           // likelihoods.jl:130
                        xs[(4 + 2 * tt + pt) * 32] = (tile_ok(Ja + tt) && t < n) ? ev : 0.0;
                    }
#pragma unroll
                for (int i = 0; i < 4; ++i) xs[(8 + i) * 32] = feed[i];
                __syncwarp();
                if (lane == 0) mbar_arrive(p2c + (xuse & 1));
            }
        }
        __syncthreads();
        // ---------------- A2: pointwise gradient ----------------
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
        const double obs_scale = (1.0 / (sigma_d * sigma_d)) * inv_b3;
        const double* yd = a.yobs + (size_t)d * n;
        double* gout = (a.grad != nullptr && cvalid) ? a.grad + chain * a.pitch + (size_t)d * n : nullptr;
        double nxa[2][2][D], nyv[2][2], nfeed[4];
        auto load_a2 = [&](int u, double (&X)[2][2][D], double (&Y)[2][2], double (&F)[4]) {
            const int Jc = 2 * u - LAGT;
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                const int J = tile_ok(Jc + tt) ? Jc + tt : -4;
#pragma unroll
                for (int dd = 0; dd < D; ++dd) ld2(xp + (size_t)dd * n, J, X[tt][0][dd], X[tt][1][dd]);
                ld2(yd, J, Y[tt][0], Y[tt][1]);
            }
            ld2(xd, 2 * u + 4, F[0], F[1]); ld2(xd, 2 * u + 5, F[2], F[3]);
        };
        load_a2(0, nxa, nyv, nfeed);
        for (int u = 0; u <= uA2_end; ++u, ++xuse) {
            const int Jc = 2 * u - LAGT;
            double xa[2][2][D], wv[2][2][D], yv[2][2], feed[4];
#pragma unroll
            for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                for (int pt = 0; pt < 2; ++pt) {
#pragma unroll
                    for (int dd = 0; dd < D; ++dd) xa[tt][pt][dd] = nxa[tt][pt][dd];
                    yv[tt][pt] = nyv[tt][pt];
                }
#pragma unroll
            for (int i = 0; i < 4; ++i) feed[i] = nfeed[i];
            if (u < uA2_end) load_a2(u + 1, nxa, nyv, nfeed);
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                const int J = tile_ok(Jc + tt) ? Jc + tt : -4;
#pragma unroll
                for (int dd = 0; dd < D; ++dd) {
                    const double* wsrc = kscr + ((size_t)(g * D + dd) * NT + (J < 0 ? 0 : J)) * 64 + lane;
                    wv[tt][0][dd] = (J < 0) ? 0.0 : wsrc[0] * inv_b1;  // likelihoods.jl:201
                    wv[tt][1][dd] = (J < 0) ? 0.0 : wsrc[32] * inv_b1;
                }
            }
            double* xs = xch + (size_t)(xuse & 1) * XS * 32;
#ifdef MAGI_DBG_WAITS
            long long w0 = clock64();
#endif
            mbar_wait(c2p + (xuse & 1), (xuse >> 1) & 1);
#ifdef MAGI_DBG_WAITS
            wc2p += clock64() - w0;
#endif
            const double cxv[2][2] = {{xs[0], xs[32]}, {xs[64], xs[96]}}, mtv[2][2] = {{xs[128], xs[160]}, {xs[192], xs[224]}};
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
#pragma unroll
                for (int pt = 0; pt < 2; ++pt) {
                    const int t = 8 * (Jc + tt) + q + 4 * pt;
                    const bool valid = tile_ok(Jc + tt) && t < n;
                    const double cx = cxv[tt][pt], mt = mtv[tt][pt];
                    const double* xv = xa[tt][pt];
                    const double* w = wv[tt][pt];
                    const double xdv = xv[d], wd = valid ? w[d] : 0.0;
                    const double y = yv[tt][pt];
                    const bool fin = valid && isfinite(y);         // likelihoods.jl:123
                    const double e0 = fin ? xdv - y : 0.0;
                    double gv = -(e0 * obs_scale);                 // likelihoods.jl:179 (e0 = 0 when the observation is missing)
                    gv -= cx * inv_b2;                             // likelihoods.jl:186
                    gv += mt * inv_b1;                             // likelihoods.jl:194
                    M::jx_col_sub(d, xv, th, w, gv);               // likelihoods.jl:214-216
                    M::jth_row_sub(d, xv, th, wd, gth);            // likelihoods.jl:219-221
                    acc_xcx += valid ? xdv * cx : 0.0;             // likelihoods.jl:150
                    acc_sse += e0 * e0;                            // likelihoods.jl:139,234
                    bad |= valid && !isfinite(gv);
                    if (valid && gout != nullptr) gout[t] = gv;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) xs[(8 + i) * 32] = feed[i];
            __syncwarp();
            if (lane == 0) mbar_arrive(p2c + (xuse & 1));
        }
        acc_xcx = quad_sum(acc_xcx);
        acc_sse = quad_sum(acc_sse);
#pragma unroll
        for (int i = 0; i < K; ++i) gth[i] = quad_sum(gth[i]);
        const unsigned badm = __ballot_sync(0xffffffffu, bad);
        if (q == 0) {
            double* r = red + ((size_t)(g * 8 + gid) * D + d) * RED;
            r[1] = acc_xcx;
            r[2] = acc_sse;
            r[3] = ((badm >> (gid * 4)) & 0xfu) ? 1.0 : 0.0;
#pragma unroll
            for (int i = 0; i < K; ++i) r[4 + i] = gth[i];
        }
    };

    if (role == 0) dispatch_dim<D>(d_rt, c_warp);
    else dispatch_dim<D>(d_rt, p_warp);
    __syncthreads();
    if (a.dbg && lane == 0) {
        long long* o = a.dbg + ((size_t)blockIdx.x * ntask + task) * 8;
        if (role == 0) { o[0] = tk1 - tk0; o[1] = tk2 - tk1; o[2] = tk3 - tk2; o[3] = clock64() - tk3; o[4] = wbar; o[5] = wp2c; o[6] = wfull; }
        else o[7] = wc2p;
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

template <int MODEL, int HB>
static cudaError_t launch_one(const BandedArgs& a, int DW, size_t smem_bytes, cudaStream_t st) {
    auto kern = banded_logpost_kernel<MODEL, HB>;
    static bool attr_set = false;    // per instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
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
