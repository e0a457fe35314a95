// Band products for models with many components (Lorenz-96, D = 64: BASELINE config 4), where the fused K1 kernels do not
// apply (their blocks keep every dimension of a chain group on one SM) and the dense route's GEMMs waste most of their
// tiles on zeros (a 128 x 128 output tile of a band of half-width 20 multiplies 176 columns for 41 useful ones).
//
//   OUT[d][c][t] = sum_t'  A_d~[t][t'] * IN(c, d, t')          A~ one of m~, C~, K~, m~^T (likelihoods.jl:129,132,133,192)
//
// is computed with the DMMA tiling of K1 (8 chains = M, 8 output times = N, contraction over 4-time chunks, NCH = 2 HB + 2
// chunks per output tile) from the fragment tables in natural order (banded_kernel.cu).  One warp owns a (dimension, pair of
// output tiles): it loads the pair's fragments ONCE into registers and sweeps over all chain groups, reading the A operand
// straight from global memory (the chain state or a [d][chain][time] plane of the previous product).  The pointwise stages
// around the products are the dense route's (dense_kernel.cu): E = f - MX, then gradient / reductions / guards.
#include <cmath>
#include "magi_internal.cuh"
#include "k1_primitives.cuh"

namespace magi {

__device__ __forceinline__ double2 ldg_frag2(const double2* p) {
    double2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

struct BandProductArgs {
    const double* frag;        // fragments of ONE view: [D][NP][NCH][32 lanes][2 tiles]
    const double* in;          // IN(c, d, t) = in[c * cs + d * ds + t]
    long long cs, ds;
    double* out;               // OUT[d][c][t] = out[d * plane + c * n + t]
    long long plane;
    int n, D, NP, n_chains;
};

template <int HB>
__global__ void __launch_bounds__(512, 1) band_product_kernel(const BandProductArgs a) {
    constexpr int NCH = 2 * HB + 2;
    const int lane = threadIdx.x & 31, gid = lane >> 2, q = lane & 3;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    const int n = a.n, NG = (a.n_chains + 7) / 8;
    const long long n_units = (long long)a.D * a.NP;
    for (long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < n_units; u += warps_total) {
        const int d = (int)(u / a.NP), p = (int)(u % a.NP);
        const double2* fr = reinterpret_cast<const double2*>(a.frag) + ((size_t)(d * a.NP + p) * NCH) * 32 + lane;
        double2 fb[NCH];
#pragma unroll
        for (int hh = 0; hh < NCH; ++hh) fb[hh] = ldg_frag2(fr + hh * 32);
        // times of this lane's A-fragment entries: chunk k covers 16p + 4k - 4HB + (0..3)
        const int t_base = 16 * p - 4 * HB + q;
#pragma unroll 1
        for (int g = 0; g < NG; ++g) {
            const int c = 8 * g + gid;
            const bool cok = c < a.n_chains;
            const double* src = a.in + (long long)(cok ? c : 0) * a.cs + (long long)d * a.ds;
            double av[NCH + 2];
#pragma unroll
            for (int k = 0; k < NCH + 2; ++k) {
                const int t = t_base + 4 * k;
                av[k] = (cok && t >= 0 && t < n) ? src[t] : 0.0;
            }
            double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
            for (int hh = 0; hh < NCH; ++hh) {
                dmma884(acc[0][0], acc[0][1], av[hh], fb[hh].x);
                dmma884(acc[1][0], acc[1][1], av[hh + 2], fb[hh].y);
            }
            if (cok) {
                double* dst = a.out + (long long)d * a.plane + (long long)c * n;
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) {
                    const int t0 = 16 * p + 8 * tt + 2 * q;
                    if (t0 < n) dst[t0] = acc[tt][0];
                    if (t0 + 1 < n) dst[t0 + 1] = acc[tt][1];
                }
            }
        }
    }
}

template <int HB>
static cudaError_t launch_bp(const BandProductArgs& a, int sm_count, cudaStream_t st) {
    const long long n_units = (long long)a.D * a.NP;
    long long blocks = (n_units + 15) / 16;
    if (blocks > sm_count) blocks = sm_count;
    band_product_kernel<HB><<<(int)blocks, 512, 0, st>>>(a);
    return cudaGetLastError();
}

// view: 0 m~, 1 C~, 2 K~, 3 m~^T of the natural-order fragment table `fragtab` ([4 views][D][NP][NCH][32][2])
cudaError_t launch_band_product(const double* fragtab, int view, const double* in, long long cs, long long ds, double* out, long long plane,
                                int n, int b, int D, int n_chains, int sm_count, cudaStream_t st) {
    const BandGeom g = band_geom(n, b);
    BandProductArgs a;
    a.NP = (g.NT + 1) / 2;
    a.frag = fragtab + (size_t)view * D * a.NP * g.NCH * 64;
    a.in = in; a.cs = cs; a.ds = ds; a.out = out; a.plane = plane; a.n = n; a.D = D; a.n_chains = n_chains;
    switch (g.HB) {
    case 0: return launch_bp<0>(a, sm_count, st);
    case 1: return launch_bp<1>(a, sm_count, st);
    case 2: return launch_bp<2>(a, sm_count, st);
    case 3: return launch_bp<3>(a, sm_count, st);
    case 4: return launch_bp<4>(a, sm_count, st);
    case 5: return launch_bp<5>(a, sm_count, st);
    case 6: return launch_bp<6>(a, sm_count, st);
    case 7: return launch_bp<7>(a, sm_count, st);
    case 8: return launch_bp<8>(a, sm_count, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace magi
