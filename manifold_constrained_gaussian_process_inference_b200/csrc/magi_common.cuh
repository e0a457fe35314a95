// Shared declarations for libmagi_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/magi_b200.h"

namespace magi {

// ---- FP64 tensor-core primitive.  On sm_100a every f64 mma.sync shape lowers to SASS DMMA.8x8x4, which issues
// once per 16 clk per SM sub-partition = 64 FMA/clk/SM, the full FP64 rate (measured 37.1 TFLOP/s,
// profiles/fp64_peaks_r01.json).  Fragment layout, lane = 4*gid + q:
//   A (8x4, row):  lane holds A[gid][q]        B (4x8, col): lane holds B[q][gid]
//   C/D (8x8):     lane holds C[gid][2q], C[gid][2q+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// cudaFuncSetAttribute acts on the CURRENT device: a process-wide "done" flag would leave the attribute unset on the second
// GPU of a process that holds handles on several devices.  One bit per device ordinal.
struct PerDeviceOnce {
    unsigned long long mask = 0;
    bool need() {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) return true;
        if ((mask >> dev) & 1ull) return false;
        mask |= 1ull << dev;
        return true;
    }
};
inline int current_sm_count() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : 148;
}

// geometry of the banded DMMA tiling (see DESIGN.md "K1"): times are tiled by 8 (one DMMA N extent) and the
// contraction runs over 4-time chunks; HB = ceil(b/4) chunks of half band on each side.
struct BandGeom {
    int n, b, HB, NCH, LAGT, WN, NT;
};
inline BandGeom band_geom(int n, int b) {
    BandGeom g;
    g.n = n; g.b = b;
    g.HB = (b + 3) / 4;
    g.NCH = 2 * g.HB + 2;
    g.LAGT = (g.HB + 1) / 2;
    g.WN = 2 * g.LAGT + 2 + g.HB;
    g.NT = (n + 7) / 8;
    return g;
}
constexpr int kMaxHB = 8;   // band half-widths up to 32 run on the windowed DMMA kernel

struct BandedArgs {
    int n, D, K, P, n_chains, NT, G, H, sigma_is_fixed, sigma_invalid, scratch_in_smem;
    long long pitch;
    const double* params;
    double* ll;
    double* grad;               // may be null (value only)
    const double* fragtab;      // [4 views][D][(NT+1)/2 pairs][NCH][32 lanes][2 tiles]
    const double* yobs;         // [D][n], non-finite = missing
    const int* nobs;            // [D]
    const double* sigma_init;   // [D]
    double beta[3];
    double inv_beta[3];
    double* scratch;            // global Ke scratch when it does not fit shared memory
    long long* dbg;             // optional per-warp phase clocks (MAGI_DBG_CLOCKS=1; null in production)
};

// arguments of the dataflow K1 (flow_kernel.cuh)
struct FlowArgs {
    int n, P, n_chains, NP, RS0, n_units, n_cblocks, sigma_is_fixed, sigma_invalid, G;
    long long pitch;
    const double* params;
    double* ll;
    double* grad;               // may be null (value only)
    const double* fragtab;      // natural order: [4 views][D][NP pairs][NCH][32 lanes][2 tiles]
    const int* units;           // wavefront order: sweep | d << 2 | pair << 8
    const double* yobs;         // [D][n]
    const int* nobs;            // [D]
    const double* sigma_init;   // [D]
    double beta[3];
    double inv_beta[3];
    long long* dbg;             // optional per-warp phase clocks (MAGI_DBG_CLOCKS=1; null in production)
};

// launches (defined in the .cu files)
cudaError_t launch_build_fragtab(const double* band_cinv, const double* band_mphi, const double* band_kinv, double* fragtab,
                                 int n, int b, int D, bool natural, double scale_c, double scale_k, cudaStream_t st);
// dataflow K1: shared-memory footprint for 16 chains (0: does not apply), row stride, wavefront unit order
size_t flow_smem_bytes(int D, int K, int KX, int n, int HB, int G, int& RS0);
std::vector<int> flow_unit_order(int D, int NP, int HB, int extra_lag);
bool model_kx(int model, int& KX);
size_t fragtab_doubles(int n, int b, int D);
size_t banded_scratch_doubles_per_cta(int G, int D, int NT);
void banded_pick_config(int model_D, int model_K, int NT, int HB, int smem_limit, int gmax, int& G, int& H, int& DW, int& scratch_in_smem, size_t& smem_bytes);
bool model_dims(int model, int& D, int& K);

}  // namespace magi
