// Batched FP64 DMMA GEMM (see gemm_f64.cuh).
#include "gemm_f64.cuh"
#include "k1_primitives.cuh"
#include <cstdlib>

namespace magi {

constexpr int GEMM_BM = 64, GEMM_BN = 64, GEMM_BK = 16, GEMM_LD = 72;

__global__ void __launch_bounds__(256) gemm_f64_dmma_kernel(const GemmArgs g) {
    __shared__ double As[2][GEMM_BK][GEMM_LD];
    __shared__ double Bs[2][GEMM_BK][GEMM_LD];
    const int m0 = blockIdx.y * GEMM_BM, n0 = blockIdx.x * GEMM_BN;
    if (g.lower_only && n0 > m0 + GEMM_BM - 1) return;
    if (g.upper_only && m0 > n0 + GEMM_BN - 1) return;
    const int z1 = blockIdx.z % g.nb1, z2 = blockIdx.z / g.nb1;
    const double* A = g.A + z1 * g.bsA1 + z2 * g.bsA2;
    const double* B = g.B + z1 * g.bsB1 + z2 * g.bsB2;
    double* C = g.C + z1 * g.bsC1 + z2 * g.bsC2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, q = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    const bool a_m_contig = (g.rsA == 1), b_n_contig = (g.csB == 1);
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    int k_begin = 0, k_end = g.K;
    if (g.k_lo_from_tile) { int t = m0 > n0 ? m0 : n0; k_begin = (t / GEMM_BK) * GEMM_BK; }
    if (g.b_lower && n0 > k_begin) k_begin = (n0 / GEMM_BK) * GEMM_BK;
    if (g.a_lower && m0 + GEMM_BM < k_end) k_end = m0 + GEMM_BM;
    if (g.a_band > 0) {
        const int lo = m0 - g.a_band, hi = m0 + GEMM_BM + g.a_band;
        if (lo > k_begin) k_begin = (lo / GEMM_BK) * GEMM_BK;
        if (hi < k_end) k_end = hi;
    }
    const int nkt = k_end > k_begin ? (k_end - k_begin + GEMM_BK - 1) / GEMM_BK : 0;
    double ra[4], rb[4];
    auto gload = [&](int kt) {
        const int k0 = k_begin + kt * GEMM_BK;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = tid + 256 * j;
            int m, k;
            if (a_m_contig) { m = idx & 63; k = idx >> 6; } else { k = idx & 15; m = idx >> 4; }
            const int gm = m0 + m, gk = k0 + k;
            ra[j] = (gm < g.M && gk < g.K) ? A[gm * g.rsA + gk * g.csA] : 0.0;
            int n, kb;
            if (b_n_contig) { n = idx & 63; kb = idx >> 6; } else { kb = idx & 15; n = idx >> 4; }
            const int gn = n0 + n, gkb = k0 + kb;
            rb[j] = (gn < g.N && gkb < g.K) ? B[gkb * g.rsB + gn * g.csB] : 0.0;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = tid + 256 * j;
            int m, k;
            if (a_m_contig) { m = idx & 63; k = idx >> 6; } else { k = idx & 15; m = idx >> 4; }
            As[buf][k][m] = ra[j];
            int n, kb;
            if (b_n_contig) { n = idx & 63; kb = idx >> 6; } else { kb = idx & 15; n = idx >> 4; }
            Bs[buf][kb][n] = rb[j];
        }
    };
    if (nkt > 0) { gload(0); sstore(0); }
    __syncthreads();
    for (int kt = 0; kt < nkt; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nkt) gload(kt + 1);
#pragma unroll
        for (int k4 = 0; k4 < GEMM_BK / 4; ++k4) {
            double a[2], b[4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) a[mi] = As[cur][k4 * 4 + q][wm * 16 + mi * 8 + gid];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[cur][k4 * 4 + q][wn * 32 + ni * 8 + gid];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
        if (kt + 1 < nkt) sstore(cur ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = m0 + wm * 16 + mi * 8 + gid, c = n0 + wn * 32 + ni * 8 + 2 * q + u;
                if (r < g.M && c < g.N) {
                    double* p = C + r * g.rsC + c * g.csC;
                    const double v = g.alpha * acc[mi][ni][u];
                    *p = (g.beta == 0.0) ? v : v + g.beta * (*p);
                }
            }
}

// ---- large-tile kernel: 128 x 128 x 16 block tiles, 16 warps (4 x 4, warp tile 32 x 32 = 4 x 4 DMMA tiles, 16 independent
// accumulate chains), 3-stage cp.async pipeline (8-byte copies with zero fill: the operands are arbitrary strided views, so
// wider copies are not generally aligned).  The stages are handed over by mbarriers instead of a block barrier per k-tile:
// full[s] (512 arrivals, each fired by the copy unit when a thread's copies of the tile have landed:
// cp.async.mbarrier.arrive.noinc) and empty[s] (one arrival per warp once it has read the tile).  A tile is requested
// BG_DIST = 1 k-tile (4096 pipe cycles) ahead into the stage that was released a whole k-tile before, so the warps of a
// block may drift apart by a tile without anybody waiting -- with one __syncthreads per k-tile the four sub-partitions
// drained together 80 times per output tile (ncu: 12 % barrier stalls, DMMA pipe 79.5 % busy).  Measured on the dense-mode
// evaluation (LV n=1281, 2048 chains, ms per evaluation): block barrier, 4 stages 2.10; mbarriers with the tile requested
// 3 / 2 / 1 tiles ahead 2.15-1.99 / 1.97 / 1.91 (requests further ahead are slower, 3 or 4 stages alike; BK = 32: 1.97).
// Each operand tile is staged in the direction it is contiguous in global memory (template flags) with a leading dimension
// that makes the DMMA fragment loads conflict-free:
//   m (n)-contiguous: S[k][m], LD = 132  -> fragment word banks 8q + 2 gid      k-contiguous: S[m][k], LD = BK + 4 -> 8 gid + 2q
// Per k4 step a warp issues 8 LDS.64 for 16 DMMAs.  Used for the dense-mode (band = n-1) operators and the large setup GEMMs.
#ifndef MAGI_GEMM_BK
#define MAGI_GEMM_BK 16
#endif
#ifndef MAGI_GEMM_ST
#define MAGI_GEMM_ST 3
#endif
#ifndef MAGI_GEMM_MBAR
#define MAGI_GEMM_MBAR 1
#endif
constexpr int BG_BM = 128, BG_BN = 128, BG_BK = MAGI_GEMM_BK, BG_ST = MAGI_GEMM_ST, BG_LDM = 132, BG_LDK = BG_BK + 4;
#ifndef MAGI_GEMM_DIST
#define MAGI_GEMM_DIST 1
#endif
constexpr int BG_DIST = MAGI_GEMM_MBAR ? MAGI_GEMM_DIST : BG_ST - 1;     // k-tiles requested ahead of the one being multiplied
constexpr int BG_CP = BG_BM * BG_BK / 512;        // 8-byte copies per thread, operand and k-tile
constexpr int BG_TILE = (BG_BK * BG_LDM > BG_BM * BG_LDK) ? BG_BK * BG_LDM : BG_BM * BG_LDK;   // doubles per operand stage
constexpr size_t BG_SMEM = sizeof(double) * 2 * BG_ST * BG_TILE + sizeof(unsigned long long) * 2 * BG_ST;
static_assert(BG_DIST >= 1 && BG_SMEM <= 227 * 1024, "pipeline depth");

// stage hand-over state of a block: the barriers live behind the operand stages; j counts the k-tiles the block has
// multiplied so far over all its output tiles / stream-K segments (tile j sits in stage j % ST, barrier phase j / ST)
struct BigPipe {
    unsigned long long* full;
    unsigned long long* empty;
    int j;
};
__device__ __forceinline__ BigPipe big_pipe_init(double* sm) {
    BigPipe p;
    p.full = reinterpret_cast<unsigned long long*>(sm + 2 * BG_ST * BG_TILE);
    p.empty = p.full + BG_ST;
    p.j = 0;
    if (MAGI_GEMM_MBAR) {
        if (threadIdx.x == 0)
            for (int s_ = 0; s_ < BG_ST; ++s_) { mbar_init(p.full + s_, 512); mbar_init(p.empty + s_, 16); }
        __syncthreads();
    }
    return p;
}

__device__ __forceinline__ void cp_async8_zfill(double* smem_dst, const double* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(d), "l"(gsrc), "r"(sz) : "memory");
}

// main loop over the k-tiles [kt_lo, kt_hi) of one 128 x 128 output tile (k-tile 0 starts at k_begin)
template <bool AK, bool BKC>
__device__ __forceinline__ void big_mainloop(const GemmArgs& g, const double* A, const double* B, int m0, int n0, int k_begin,
                                             int kt_lo, int kt_hi, double* As, double* Bs, BigPipe& pipe, double (&acc)[4][4][2]) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, q = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    // this thread's BG_CP + BG_CP copies per k-tile: running global pointers, per-copy steps and row masks are set up once
    // (k-contiguous operand: copy j covers row am + (512 / BK) j at k = ak; otherwise row am at k = ak + 4 j)
    constexpr int RPT = 512 / BG_BK;                   // rows per pass of a k-contiguous operand
    const int am = AK ? tid / BG_BK : tid & 127, ak = AK ? tid % BG_BK : tid >> 7;
    const int bn = BKC ? tid / BG_BK : tid & 127, bk = BKC ? tid % BG_BK : tid >> 7;
    const double* ap = A + (long long)(m0 + am) * g.rsA + (long long)(k_begin + kt_lo * BG_BK + ak) * g.csA;
    const double* bp = B + (long long)(n0 + bn) * g.csB + (long long)(k_begin + kt_lo * BG_BK + bk) * g.rsB;
    const long long a_step = AK ? (long long)RPT * g.rsA : 4 * g.csA, b_step = BKC ? (long long)RPT * g.csB : 4 * g.rsB;
    const long long a_kt = (long long)BG_BK * g.csA, b_kt = (long long)BG_BK * g.rsB;
    unsigned a_mask = 0, b_mask = 0;                   // bit j: the row of copy j exists
#pragma unroll
    for (int j = 0; j < BG_CP; ++j) {
        a_mask |= ((m0 + am + (AK ? RPT * j : 0)) < g.M ? 1u : 0u) << j;
        b_mask |= ((n0 + bn + (BKC ? RPT * j : 0)) < g.N ? 1u : 0u) << j;
    }
    const int a_soff = AK ? am * BG_LDK + ak : ak * BG_LDM + am, b_soff = BKC ? bn * BG_LDK + bk : bk * BG_LDM + bn;
    constexpr int a_sstep = AK ? RPT * BG_LDK : 4 * BG_LDM, b_sstep = BKC ? RPT * BG_LDK : 4 * BG_LDM;
    int k_left_a = g.K - (k_begin + kt_lo * BG_BK + ak), k_left_b = g.K - (k_begin + kt_lo * BG_BK + bk);   // copy j is inside K iff its k offset < k_left
    auto load_tile = [&](int kt) {
        if (kt < kt_hi) {
            const int j = pipe.j + (kt - kt_lo), stg = j % BG_ST;
            if (MAGI_GEMM_MBAR && j >= BG_ST) mbar_wait(pipe.empty + stg, (unsigned)((j / BG_ST) - 1) & 1u);   // tile j - ST has been read by all warps
            double* as = As + stg * BG_TILE + a_soff;
            double* bs = Bs + stg * BG_TILE + b_soff;
#pragma unroll
            for (int j_ = 0; j_ < BG_CP; ++j_) {
                const bool ok = ((a_mask >> j_) & 1u) && (AK ? 0 : 4 * j_) < k_left_a;
                cp_async8_zfill(as + j_ * a_sstep, ok ? ap + j_ * a_step : A, ok);
                const bool okb = ((b_mask >> j_) & 1u) && (BKC ? 0 : 4 * j_) < k_left_b;
                cp_async8_zfill(bs + j_ * b_sstep, okb ? bp + j_ * b_step : B, okb);
            }
            ap += a_kt; bp += b_kt; k_left_a -= BG_BK; k_left_b -= BG_BK;
        }
        if (MAGI_GEMM_MBAR) {
            // fires when this thread's copies above have landed, wherever the thread is by then (no-op group when kt >= kt_hi: nobody waits for it)
            if (kt < kt_hi) asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" :: "r"(smem_u32(pipe.full + (pipe.j + (kt - kt_lo)) % BG_ST)) : "memory");
        } else {
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
    };
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
    for (int s_ = 0; s_ < BG_DIST; ++s_) load_tile(kt_lo + s_);
    for (int kt = kt_lo; kt < kt_hi; ++kt) {
        const int j = pipe.j + (kt - kt_lo), stg = j % BG_ST;
        if (MAGI_GEMM_MBAR) {
            load_tile(kt + BG_DIST);
            mbar_wait(pipe.full + stg, (unsigned)(j / BG_ST) & 1u);                     // the copies of all 512 threads for tile kt have landed
        } else {
            asm volatile("cp.async.wait_group %0;\n" :: "n"(BG_ST - 2) : "memory");
            __syncthreads();
            load_tile(kt + BG_DIST);
        }
        const double* as = As + stg * BG_TILE;
        const double* bs = Bs + stg * BG_TILE;
#pragma unroll
        for (int k4 = 0; k4 < BG_BK / 4; ++k4) {
            double af[4], bf[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const int m = wm * 32 + mi * 8 + gid, k = k4 * 4 + q;
                af[mi] = AK ? as[m * BG_LDK + k] : as[k * BG_LDM + m];
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int n = wn * 32 + ni * 8 + gid, k = k4 * 4 + q;
                bf[ni] = BKC ? bs[n * BG_LDK + k] : bs[k * BG_LDM + n];
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
        if (MAGI_GEMM_MBAR) {
            __syncwarp();
            if (lane == 0) mbar_arrive(pipe.empty + stg);
        }
    }
    if (!MAGI_GEMM_MBAR) {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();          // the stages may be refilled by the next segment
    }
    pipe.j += kt_hi - kt_lo;
}

__device__ __forceinline__ void big_store(const GemmArgs& g, double* C, int m0, int n0, const double (&acc)[4][4][2]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = m0 + wm * 32 + mi * 8 + gid, c = n0 + wn * 32 + ni * 8 + 2 * q + u;
                if (r < g.M && c < g.N) {
                    double* p = C + (long long)r * g.rsC + (long long)c * g.csC;
                    const double v = g.alpha * acc[mi][ni][u];
                    *p = (g.beta == 0.0) ? v : v + g.beta * (*p);
                }
            }
}

template <bool AK, bool BKC>
__global__ void __launch_bounds__(512, 1) gemm_f64_dmma_big_kernel(const GemmArgs g) {
    extern __shared__ __align__(16) double sm[];
    double* As = sm;                                   // [ST][BG_TILE]
    double* Bs = sm + BG_ST * BG_TILE;
    const int m0 = blockIdx.y * BG_BM, n0 = blockIdx.x * BG_BN;
    if (g.lower_only && n0 > m0 + BG_BM - 1) return;
    if (g.upper_only && m0 > n0 + BG_BN - 1) return;
    const int z1 = blockIdx.z % g.nb1, z2 = blockIdx.z / g.nb1;
    const double* A = g.A + z1 * g.bsA1 + z2 * g.bsA2;
    const double* B = g.B + z1 * g.bsB1 + z2 * g.bsB2;
    double* C = g.C + z1 * g.bsC1 + z2 * g.bsC2;
    int k_begin = 0, k_end = g.K;
    if (g.k_lo_from_tile) { int t = m0 > n0 ? m0 : n0; k_begin = (t / BG_BK) * BG_BK; }
    if (g.b_lower && n0 > k_begin) k_begin = (n0 / BG_BK) * BG_BK;
    if (g.a_lower && m0 + BG_BM < k_end) k_end = m0 + BG_BM;
    if (g.a_band > 0) {
        const int lo = m0 - g.a_band, hi = m0 + BG_BM + g.a_band;
        if (lo > k_begin) k_begin = (lo / BG_BK) * BG_BK;
        if (hi < k_end) k_end = hi;
    }
    const int nkt = k_end > k_begin ? (k_end - k_begin + BG_BK - 1) / BG_BK : 0;
    double acc[4][4][2];
    BigPipe pipe = big_pipe_init(sm);
    big_mainloop<AK, BKC>(g, A, B, m0, n0, k_begin, 0, nkt, As, Bs, pipe, acc);
    big_store(g, C, m0, n0, acc);
}

// ---- stream-K variant: a persistent grid of one block per SM; the (tile, k-tile) work units are dealt out in equal contiguous
// ranges, so the load is balanced whatever the tile count (LV n=1281 x 2048 chains x 2 dimensions is 320 tiles = 2.16 waves of
// 148).  A block's range starts inside a tile (leading partial: accumulators -> work space, flag released), covers whole
// tiles, and ends inside a tile (trailing partial: it waits for the NEXT block's leading partial -- written at that block's
// very start -- adds it and stores).  With at least as many tiles as blocks a tile is shared by at most two blocks and the
// sum order (low k range + high k range) is fixed: results are deterministic.
template <bool AK, bool BKC>
__global__ void __launch_bounds__(512, 1) gemm_f64_dmma_streamk_kernel(const GemmArgs g, int tiles_m, int tiles_n, int batch) {
    extern __shared__ __align__(16) double sm[];
    double* As = sm;
    double* Bs = sm + BG_ST * BG_TILE;
    const int nkt = (g.K + BG_BK - 1) / BG_BK;
    const long long units = (long long)tiles_m * tiles_n * batch * nkt;
    long long u = units * blockIdx.x / gridDim.x;
    const long long u_end = units * (blockIdx.x + 1) / gridDim.x;
    double acc[4][4][2];
    BigPipe pipe = big_pipe_init(sm);
    while (u < u_end) {
        const long long tile = u / nkt;
        const int kt0 = (int)(u - tile * nkt);
        const int kt1 = (int)((u_end - u) < (long long)(nkt - kt0) ? kt0 + (u_end - u) : nkt);
        const int tn = (int)(tile % tiles_n), tm = (int)((tile / tiles_n) % tiles_m), z = (int)(tile / ((long long)tiles_n * tiles_m));
        const int z1 = z % g.nb1, z2 = z / g.nb1;
        const double* A = g.A + z1 * g.bsA1 + z2 * g.bsA2;
        const double* B = g.B + z1 * g.bsB1 + z2 * g.bsB2;
        double* C = g.C + z1 * g.bsC1 + z2 * g.bsC2;
        const int m0 = tm * BG_BM, n0 = tn * BG_BN;
        big_mainloop<AK, BKC>(g, A, B, m0, n0, 0, kt0, kt1, As, Bs, pipe, acc);
        if (kt0 > 0) {                                        // leading partial of this block: hand the accumulators to the owner
            double* w = g.sk_work + (size_t)blockIdx.x * (BG_BM * BG_BN);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                    for (int e = 0; e < 2; ++e) w[(size_t)((mi * 4 + ni) * 2 + e) * 512 + threadIdx.x] = acc[mi][ni][e];
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;\n" :: "l"(g.sk_flags + blockIdx.x), "r"(g.sk_epoch) : "memory");
        } else if (kt1 < nkt) {                               // trailing partial: this block owns the tile
            if (threadIdx.x == 0) {
                unsigned v = 0;
                do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(g.sk_flags + blockIdx.x + 1) : "memory"); } while (v != g.sk_epoch);
            }
            __syncthreads();
            const double* w = g.sk_work + (size_t)(blockIdx.x + 1) * (BG_BM * BG_BN);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                    for (int e = 0; e < 2; ++e) acc[mi][ni][e] += __ldcg(w + (size_t)((mi * 4 + ni) * 2 + e) * 512 + threadIdx.x);
            big_store(g, C, m0, n0, acc);
        } else {
            big_store(g, C, m0, n0, acc);
        }
        u += kt1 - kt0;
    }
}

template <bool AK, bool BKC>
static cudaError_t launch_big(const GemmArgs& g, int batch, cudaStream_t st) {
    auto kern = gemm_f64_dmma_big_kernel<AK, BKC>;
    const size_t smem = BG_SMEM;
    static PerDeviceOnce once;       // per instantiation
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((g.N + BG_BN - 1) / BG_BN, (g.M + BG_BM - 1) / BG_BM, batch);
    kern<<<grid, 512, smem, st>>>(g);
    return cudaGetLastError();
}

template <bool AK, bool BKC>
static cudaError_t launch_streamk(const GemmArgs& g, int tiles_m, int tiles_n, int batch, int blocks, cudaStream_t st) {
    auto kern = gemm_f64_dmma_streamk_kernel<AK, BKC>;
    const size_t smem = BG_SMEM;
    static PerDeviceOnce once;       // per instantiation
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<blocks, 512, smem, st>>>(g, tiles_m, tiles_n, batch);
    return cudaGetLastError();
}

// ---- skinny remainder: a handful of output rows (M <= 16, e.g. the 1281st row next to ten 128-row tiles).  B is read exactly
// once: a block owns 16 output columns, two per warp, lanes stride over k (coalesced when B is k-contiguous, several loads in
// flight per lane); the A rows, four at a time, are staged in shared memory in k-chunks, so their layout does not matter.
constexpr int ROWS_KC = 1024, ROWS_COLS = 16;
__global__ void __launch_bounds__(256) gemm_f64_rows_kernel(const GemmArgs g) {
    __shared__ double As[4][ROWS_KC];
    const int z1 = blockIdx.y % g.nb1, z2 = blockIdx.y / g.nb1;
    const double* A = g.A + z1 * g.bsA1 + z2 * g.bsA2;
    const double* B = g.B + z1 * g.bsB1 + z2 * g.bsB2;
    double* C = g.C + z1 * g.bsC1 + z2 * g.bsC2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c0 = blockIdx.x * ROWS_COLS + 2 * warp;
    const bool v0 = c0 < g.N, v1 = c0 + 1 < g.N;
    for (int r0 = 0; r0 < g.M; r0 += 4) {
        double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
        for (int kc = 0; kc < g.K; kc += ROWS_KC) {
            const int kn = g.K - kc < ROWS_KC ? g.K - kc : ROWS_KC;
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i)
                for (int k = threadIdx.x; k < kn; k += 256)
                    As[i][k] = (r0 + i < g.M) ? A[(long long)(r0 + i) * g.rsA + (long long)(kc + k) * g.csA] : 0.0;
            __syncthreads();
            const double* b0 = B + (long long)c0 * g.csB + (long long)kc * g.rsB;
            const double* b1 = b0 + g.csB;
#pragma unroll 4
            for (int k = lane; k < kn; k += 32) {
                const double x0 = v0 ? b0[(long long)k * g.rsB] : 0.0, x1 = v1 ? b1[(long long)k * g.rsB] : 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double av = As[i][k];
                    acc[i][0] += av * x0;
                    acc[i][1] += av * x1;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                double v = acc[i][u];
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && r0 + i < g.M && c0 + u < g.N) {
                    double* p = C + (long long)(r0 + i) * g.rsC + (long long)(c0 + u) * g.csC;
                    v *= g.alpha;
                    *p = (g.beta == 0.0) ? v : v + g.beta * (*p);
                }
            }
    }
}

static cudaError_t launch_small(const GemmArgs& g, int batch, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0) return cudaSuccess;
    dim3 grid((g.N + GEMM_BN - 1) / GEMM_BN, (g.M + GEMM_BM - 1) / GEMM_BM, batch);
    gemm_f64_dmma_kernel<<<grid, 256, 0, st>>>(g);
    return cudaGetLastError();
}

cudaError_t launch_gemm(const GemmArgs& g, int batch, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0 || batch <= 0) return cudaSuccess;
    static const bool force_small = getenv("MAGI_GEMM_SMALL") != nullptr;
    static const bool no_streamk = getenv("MAGI_GEMM_NO_STREAMK") != nullptr;
    const int sm_count = g.sm_count > 0 ? g.sm_count : current_sm_count();      // of the device the caller launches on
    const bool a_ok = (g.rsA == 1 || g.csA == 1), b_ok = (g.rsB == 1 || g.csB == 1);
    const bool ak = (g.rsA != 1);          // A(m, k): unit stride along k unless rsA == 1 (m-contiguous)
    const bool bkc = (g.csB != 1);         // B(k, n): unit stride along k unless csB == 1 (n-contiguous)
    if (!force_small && a_ok && b_ok && g.K >= 64 && g.M >= 96 && g.N >= 96) {
        // stream-K over 128 x 128 tiles (edge tiles are predicated); a remainder of at most 16 rows is cut off and handled
        // by the skinny kernel instead of paying a whole tile row for it
        const int m_rem = g.M % BG_BM;
        const bool cut = (m_rem > 0 && m_rem <= 16 && g.M > BG_BM);
        const int tm = cut ? g.M / BG_BM : (g.M + BG_BM - 1) / BG_BM, tn = (g.N + BG_BN - 1) / BG_BN;
        if (!no_streamk && g.sk_work && g.sk_flags && !g.lower_only && !g.k_lo_from_tile && !g.upper_only && !g.a_lower && !g.b_lower && g.a_band == 0 && sm_count <= kStreamKSlots - 1 &&
            (long long)tm * tn * batch >= sm_count) {
            GemmArgs m = g; if (cut) m.M = tm * BG_BM;
            cudaError_t e = ak ? (bkc ? launch_streamk<true, true>(m, tm, tn, batch, sm_count, st) : launch_streamk<true, false>(m, tm, tn, batch, sm_count, st))
                               : (bkc ? launch_streamk<false, true>(m, tm, tn, batch, sm_count, st) : launch_streamk<false, false>(m, tm, tn, batch, sm_count, st));
            if (e != cudaSuccess) return e;
            if (cut) {                               // bottom rows [m.M, M), all columns
                GemmArgs r = g; r.A = g.A + (long long)m.M * g.rsA; r.C = g.C + (long long)m.M * g.rsC; r.M = g.M - m.M;
                gemm_f64_rows_kernel<<<dim3((g.N + ROWS_COLS - 1) / ROWS_COLS, batch), 256, 0, st>>>(r);
                e = cudaGetLastError();
                if (e != cudaSuccess) return e;
                if (g.extra_launches) ++*g.extra_launches;
            }
            return cudaSuccess;
        }
        const long long tiles = (long long)((g.N + BG_BN - 1) / BG_BN) * ((g.M + BG_BM - 1) / BG_BM) * batch;
        if (tiles >= 96) {
            if (ak) return bkc ? launch_big<true, true>(g, batch, st) : launch_big<true, false>(g, batch, st);
            return bkc ? launch_big<false, true>(g, batch, st) : launch_big<false, false>(g, batch, st);
        }
    }
    return launch_small(g, batch, st);
}

}  // namespace magi
