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
#include <type_traits>
#pragma once
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

// ---- async-copy helpers ----
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(sa), "l"(gsrc) : "memory");
}
// 8-byte async copy of one state value into this lane's private staging slot; zero-fills when !ok (src-size 0)
__device__ __forceinline__ void cp_async8_zfill(double* smem_dst, const double* gsrc, bool ok) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = ok ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(sa), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }
__device__ __forceinline__ void named_barrier(int id, int nthreads) { asm volatile("bar.sync %0, %1;\n" :: "r"(id), "r"(nthreads) : "memory"); }

// One block = G chain-groups (8 chains each) x DW dimension slots, one warp per (group, dim) task, sweeping the time axis
// TWO 8-time tiles per step (halves the per-step overhead -- window shifts, ring bookkeeping, barriers, address math -- and
// gives the scheduler four independent DMMA accumulate chains plus two independent pointwise evaluations per step).
//   A1: mx = m~ x_d, e = f - mx (sliding register windows), Ke = K~ e -> Ke scratch (shared memory); sum e.Ke
//   A2: Cx = C~ x_d ; mt = m~^T Ke_d ; pointwise gradient incl. the ODE Jacobian terms (need Ke of all dimensions)
// The four table-fragment blocks a step needs are staged ONCE per dimension in a double-buffered cp.async ring in shared
// memory, shared by the G warps of that dimension (one named barrier per step), and read with 16-byte LDS (two chunks per
// load): each fragment block leaves L2 once per block.
// calls f(std::integral_constant<int, d>) for the runtime (warp-uniform) dimension d: the model functors then see a
// compile-time dimension and compile to straight-line code
template <int D, class F> __device__ __forceinline__ void dispatch_dim(int d, F& f) {
    if constexpr (D >= 1) { if (d == 0) { f(std::integral_constant<int, 0>{}); return; } }
    if constexpr (D >= 2) { if (d == 1) { f(std::integral_constant<int, 1>{}); return; } }
    if constexpr (D >= 3) { if (d == 2) { f(std::integral_constant<int, 2>{}); return; } }
    if constexpr (D >= 4) { if (d == 3) { f(std::integral_constant<int, 3>{}); return; } }
    if constexpr (D >= 5) { if (d == 4) { f(std::integral_constant<int, 4>{}); return; } }
}

// block = G chain-groups x D dimension warps (at most 16 warps): the register budget follows from the model's D
// (FN: 8 warps = 256 threads -> up to 255 registers per thread)
template <int MODEL> constexpr int banded_max_threads() { return Ode<MODEL>::D * 128 > 512 ? 512 : Ode<MODEL>::D * 128; }

template <int MODEL, int HB>
__global__ void __launch_bounds__(banded_max_threads<MODEL>(), 1) banded_logpost_kernel(const BandedArgs a) {
    using M = Ode<MODEL>;
    constexpr int D = M::D, K = M::K;
    constexpr int NCH = 2 * HB + 2, LAGT = (HB + 1) / 2, WN = 2 * LAGT + 2 + HB, W2 = WN + 2;
    constexpr int RED = 4 + K;   // e.Ke, x.Cx, sse, bad flag, theta-gradient partials
    constexpr int BLK = NCH * 32;               // doubles per fragment block (one view, one tile)
    extern __shared__ __align__(128) double smem[];
    const int NT = a.NT, n = a.n, G = a.G;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
    const int nwarps = blockDim.x >> 5;
    const int DW = nwarps / G;                       // dimensions processed concurrently by the block
    const int g = warp % G, dslot = warp / G;
    const size_t scr_doubles = (size_t)G * D * NT * 64;
    // shared memory: [Ke scratch (if it fits)] [rings: DW x 2 stages x 4 blocks] [red]
    double* kscr = a.scratch_in_smem ? smem : a.scratch + (size_t)blockIdx.x * scr_doubles;
    double* rings = smem + (a.scratch_in_smem ? scr_doubles : 0);
    double* ring = rings + (size_t)dslot * 2 * 4 * BLK;
    unsigned long long* mbars = reinterpret_cast<unsigned long long*>(rings + (size_t)DW * 2 * 4 * BLK);   // [DW][2 stages]
    unsigned long long* full = mbars + dslot * 2;
    double* red = rings + (size_t)DW * 2 * 4 * BLK + 2 * 8;                                  // [G*8][D][RED] (after 16 mbarrier slots)
    const int ring_threads = G * 32;
    const bool ring_leader = (g == 0 && lane == 0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < DW * 2; ++i) mbar_init(mbars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    unsigned ring_use = 0;                           // number of ring stages consumed so far by this ring (parity tracking)

    const long long chain = (long long)blockIdx.x * (G * 8) + g * 8 + gid;
    const bool cvalid = chain < a.n_chains;
    const double* xp = a.params + (cvalid ? chain : (long long)a.n_chains - 1) * a.pitch;
    double th[M::KX];
#pragma unroll
    for (int i = 0; i < K; ++i) th[i] = xp[(size_t)n * D + i];
    M::prepare(th);
    const double inv_b1 = a.inv_beta[0], inv_b2 = a.inv_beta[1], inv_b3 = a.inv_beta[2];
    long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0;
#ifdef MAGI_DBG_FINE
    long long fine[4] = {0, 0, 0, 0};
#endif
    if (a.dbg) tk0 = clock64();

    // the ring leader stages up to four fragment blocks of a step with one TMA bulk copy each; `use` is the ring-use index
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

    // state loads for tile J (times 8J+q, 8J+q+4); zero outside [0, n)
    auto ld2 = [&](const double* base, int J, double& v0, double& v1) {
        const int t0 = 8 * J + q, t1 = t0 + 4;
        v0 = (t0 >= 0 && t0 < n) ? base[t0] : 0.0;
        v1 = (t1 >= 0 && t1 < n) ? base[t1] : 0.0;
    };

    // ---------------- A1: mx, e, Ke ----------------
    // step u: pushes x tiles (2u, 2u+1); mx / e for tiles (Ja, Ja+1), Ja = 2u - LAGT; Ke for tiles (Jb, Jb+1),
    // Jb = 2u - 2 - 2 LAGT, from the e window as left by step u-1.
    auto phase_a1 = [&](auto dconst) {
        constexpr int d = decltype(dconst)::value;
        double xw[W2], ew[W2];
#pragma unroll
        for (int i = 0; i < W2; ++i) { xw[i] = 0.0; ew[i] = 0.0; }
        double acc_eke = 0.0;
        const double* xd = xp + (size_t)d * n;
        double* ks = kscr + ((size_t)(g * D + d) * NT) * 64 + lane;
        const double* ft0 = a.fragtab + ((size_t)(0 * D + d) * NT) * BLK;
        const double* ft2 = a.fragtab + ((size_t)(2 * D + d) * NT) * BLK;
        const int u_end = (NT + 2 * LAGT + 2) / 2;            // last step with a valid Ke tile: Jb + 1 >= NT - 1
        auto stage_step = [&](int u, unsigned use) {        // fragments needed by step u
            if (u <= u_end) {
                const int Ja = 2 * u - LAGT, Jb = 2 * u - 2 - 2 * LAGT;
                stage_blocks(use, tile_ok(Ja) ? ft0 + (size_t)Ja * BLK : nullptr, tile_ok(Ja + 1) ? ft0 + (size_t)(Ja + 1) * BLK : nullptr,
                             tile_ok(Jb) ? ft2 + (size_t)Jb * BLK : nullptr, tile_ok(Jb + 1) ? ft2 + (size_t)(Jb + 1) * BLK : nullptr);
            }
        };
        named_barrier(1 + dslot, ring_threads);              // every warp of the ring is done with the previous phase's stages
        stage_step(0, ring_use);
        double nx[4];
        ld2(xd, 0, nx[0], nx[1]); ld2(xd, 1, nx[2], nx[3]);
        for (int u = 0; u <= u_end; ++u, ++ring_use) {
            named_barrier(1 + dslot, ring_threads);          // all warps finished step u-1: its stage may be overwritten
            stage_step(u + 1, ring_use + 1);
            mbar_wait(full + (ring_use & 1), (ring_use >> 1) & 1);
            const double2* fr = reinterpret_cast<const double2*>(ring + (size_t)(ring_use & 1) * 4 * BLK) + lane;
#pragma unroll
            for (int i = 0; i < W2 - 4; ++i) xw[i] = xw[i + 4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xw[W2 - 4 + i] = nx[i];
            ld2(xd, 2 * u + 2, nx[0], nx[1]); ld2(xd, 2 * u + 3, nx[2], nx[3]);       // next step's window feed
            const int Ja = 2 * u - LAGT, Jb = 2 * u - 2 - 2 * LAGT;
            const bool va = tile_ok(Ja) || tile_ok(Ja + 1), vb = tile_ok(Jb) || tile_ok(Jb + 1);
            double xa[2][2][D];                                // [tile][point][dim]; zeros where the tile is outside the grid
#pragma unroll
            for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                for (int dd = 0; dd < D; ++dd) {
                    if (dd == d) { xa[tt][0][dd] = xw[HB + 2 * tt]; xa[tt][1][dd] = xw[HB + 2 * tt + 1]; }
                    else ld2(xp + (size_t)dd * n, tile_ok(Ja + tt) ? Ja + tt : -4, xa[tt][0][dd], xa[tt][1][dd]);
                }
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
            // branch-free: the four points are independent dependent-chains the scheduler can interleave
            double e[2][2];
#pragma unroll
            for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                for (int pt = 0; pt < 2; ++pt) {
                    const int t = 8 * (Ja + tt) + q + 4 * pt;
                    const double ev = M::f(d, xa[tt][pt], th) - m[tt][pt];           // likelihoods.jl:130
                    e[tt][pt] = (va && tile_ok(Ja + tt) && t < n) ? ev : 0.0;
                }
#pragma unroll
            for (int i = 0; i < W2 - 4; ++i) ew[i] = ew[i + 4];
            ew[W2 - 4] = e[0][0]; ew[W2 - 3] = e[0][1]; ew[W2 - 2] = e[1][0]; ew[W2 - 1] = e[1][1];
        }
        acc_eke = quad_sum(acc_eke);
        if (q == 0) red[((size_t)(g * 8 + gid) * D + d) * RED + 0] = acc_eke;
    };
#ifdef MAGI_STAGGER
    if (dslot & 1) { const long long t0 = clock64(); while (clock64() - t0 < MAGI_STAGGER) {} }
#endif
    for (int dr = dslot; dr < D; dr += DW) dispatch_dim<D>(dr, phase_a1);
    if (a.dbg) tk1 = clock64();
    __syncthreads();
    if (a.dbg) tk2 = clock64();

    // ---------------- A2: Cx, m^T Ke, pointwise gradient ----------------
    // step u: pushes x and Ke tiles (2u, 2u+1); outputs for tiles (Jc, Jc+1), Jc = 2u - LAGT
    auto phase_a2 = [&](auto dconst) {
        constexpr int d = decltype(dconst)::value;
        double xw[W2], kw[W2];
#pragma unroll
        for (int i = 0; i < W2; ++i) { xw[i] = 0.0; kw[i] = 0.0; }
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
        const double* ft1 = a.fragtab + ((size_t)(1 * D + d) * NT) * BLK;
        const double* ft3 = a.fragtab + ((size_t)(3 * D + d) * NT) * BLK;
        double* gout = (a.grad != nullptr && cvalid) ? a.grad + chain * a.pitch + (size_t)d * n : nullptr;
        const int u_end = (NT - 1 + LAGT) / 2;
        auto stage_step = [&](int u, unsigned use) {
            if (u <= u_end) {
                const int Jc = 2 * u - LAGT;
                stage_blocks(use, tile_ok(Jc) ? ft1 + (size_t)Jc * BLK : nullptr, tile_ok(Jc + 1) ? ft1 + (size_t)(Jc + 1) * BLK : nullptr,
                             tile_ok(Jc) ? ft3 + (size_t)Jc * BLK : nullptr, tile_ok(Jc + 1) ? ft3 + (size_t)(Jc + 1) * BLK : nullptr);
            }
        };
        named_barrier(1 + dslot, ring_threads);
        stage_step(0, ring_use);
        double nx[4];
        ld2(xd, 0, nx[0], nx[1]); ld2(xd, 1, nx[2], nx[3]);
        for (int u = 0; u <= u_end; ++u, ++ring_use) {
#ifdef MAGI_DBG_FINE
            const long long f0 = clock64();
#endif
            named_barrier(1 + dslot, ring_threads);
            stage_step(u + 1, ring_use + 1);
            mbar_wait(full + (ring_use & 1), (ring_use >> 1) & 1);
#ifdef MAGI_DBG_FINE
            const long long f1b = clock64();
            fine[0] += f1b - f0;
#endif
            const double2* fr = reinterpret_cast<const double2*>(ring + (size_t)(ring_use & 1) * 4 * BLK) + lane;
#pragma unroll
            for (int i = 0; i < W2 - 4; ++i) { xw[i] = xw[i + 4]; kw[i] = kw[i + 4]; }
#pragma unroll
            for (int i = 0; i < 4; ++i) xw[W2 - 4 + i] = nx[i];
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                const bool ok = tile_ok(2 * u + tt);
                kw[W2 - 4 + 2 * tt] = ok ? ks[(size_t)(2 * u + tt) * 64] : 0.0;
                kw[W2 - 3 + 2 * tt] = ok ? ks[(size_t)(2 * u + tt) * 64 + 32] : 0.0;
            }
            ld2(xd, 2 * u + 2, nx[0], nx[1]); ld2(xd, 2 * u + 3, nx[2], nx[3]);
            const int Jc = 2 * u - LAGT;
            if (tile_ok(Jc) || tile_ok(Jc + 1)) {
                double xa[2][2][D], wv[2][2][D], yv[2][2];
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) {
                    const int J = tile_ok(Jc + tt) ? Jc + tt : -4;
#pragma unroll
                    for (int dd = 0; dd < D; ++dd) {
                        if (dd == d) { xa[tt][0][dd] = xw[HB + 2 * tt]; xa[tt][1][dd] = xw[HB + 2 * tt + 1]; }
                        else ld2(xp + (size_t)dd * n, J, xa[tt][0][dd], xa[tt][1][dd]);
                        const double* wsrc = kscr + ((size_t)(g * D + dd) * NT + (J < 0 ? 0 : J)) * 64 + lane;
                        wv[tt][0][dd] = (J < 0) ? 0.0 : wsrc[0] * inv_b1;  // likelihoods.jl:201
                        wv[tt][1][dd] = (J < 0) ? 0.0 : wsrc[32] * inv_b1;
                    }
                    ld2(yd, J, yv[tt][0], yv[tt][1]);
                }
                double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, um[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#ifdef MAGI_DBG_FINE
                const long long f2 = clock64() + (long long)(xa[0][0][0] * 0.0) + (long long)(wv[1][1][D - 1] * 0.0) + (long long)(yv[1][1] * 0.0);
#endif
#pragma unroll
                for (int hp = 0; hp < NCH / 2; ++hp) {
                    const double2 fa0 = fr[hp * 32], fa1 = fr[BLK / 2 + hp * 32], fb0 = fr[BLK + hp * 32], fb1 = fr[3 * BLK / 2 + hp * 32];
                    dmma884(c[0][0], c[0][1], xw[2 * hp], fa0.x);          dmma884(c[1][0], c[1][1], xw[2 * hp + 2], fa1.x);      // likelihoods.jl:133
                    dmma884(um[0][0], um[0][1], kw[2 * hp], fb0.x);        dmma884(um[1][0], um[1][1], kw[2 * hp + 2], fb1.x);    // likelihoods.jl:192
                    dmma884(c[0][0], c[0][1], xw[2 * hp + 1], fa0.y);      dmma884(c[1][0], c[1][1], xw[2 * hp + 3], fa1.y);
                    dmma884(um[0][0], um[0][1], kw[2 * hp + 1], fb0.y);    dmma884(um[1][0], um[1][1], kw[2 * hp + 3], fb1.y);
                }
#ifdef MAGI_DBG_FINE
                const long long f3 = clock64() + (long long)(c[0][0] * 0.0) + (long long)(um[1][1] * 0.0) + (long long)(c[1][0] * 0.0) + (long long)(um[0][1] * 0.0);
#endif
                // branch-free pointwise stage: four independent points per step (invalid points carry x = w = 0)
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) {
#pragma unroll
                    for (int pt = 0; pt < 2; ++pt) {
                        const int t = 8 * (Jc + tt) + q + 4 * pt;
                        const bool valid = tile_ok(Jc + tt) && t < n;
                        const double cx = c[tt][pt], mt = um[tt][pt];
                        const double* xv = xa[tt][pt];
                        const double* w = wv[tt][pt];
                        const double xdv = xv[d], wd = valid ? w[d] : 0.0;
                        const double y = yv[tt][pt];
                        const bool fin = valid && isfinite(y);         // likelihoods.jl:123
                        const double e0 = fin ? xdv - y : 0.0;
                        double gv = -(e0 * inv_sig2) * inv_b3;         // likelihoods.jl:179 (e0 = 0 when the observation is missing)
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
#ifdef MAGI_DBG_FINE
                const long long f4 = clock64() + (long long)(acc_xcx * 0.0) + (long long)(gth[K - 1] * 0.0);
                fine[1] += f2 - f1b; fine[2] += f3 - f2; fine[3] += f4 - f3;
#endif
            }
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
#ifdef MAGI_STAGGER
    if (dslot & 1) { const long long t0 = clock64(); while (clock64() - t0 < MAGI_STAGGER) {} }
#endif
    for (int dr = dslot; dr < D; dr += DW) dispatch_dim<D>(dr, phase_a2);
    if (a.dbg) tk3 = clock64();
    __syncthreads();
    if (a.dbg && lane == 0) {
        long long* o = a.dbg + ((size_t)blockIdx.x * nwarps + warp) * 8;
        o[0] = tk1 - tk0; o[1] = tk2 - tk1; o[2] = tk3 - tk2; o[3] = clock64() - tk3; o[4] = 0; o[5] = 0; o[6] = 0; o[7] = 0;
#ifdef MAGI_DBG_FINE
        o[4] = fine[0]; o[5] = fine[1]; o[6] = fine[2]; o[7] = fine[3];
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
    const int threads = a.G * DW * a.H * 32;
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
