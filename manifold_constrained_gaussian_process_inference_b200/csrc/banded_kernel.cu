// K6 (band -> fragment tables) and the host-side configuration / dispatch of K1 (kernel: banded_kernel.cuh).
//
// Replaces, for a batch of independent chains, the reference's
//   log_likelihood_and_gradient_banded          src/likelihoods.jl:43-257
//   LogDensityProblems.logdensity_and_gradient  src/logdensityproblems_interface.jl:176-267
//
// The kernel, its formulation (8 chains x 8 output times per DMMA, contraction over 4-time chunks, sliding register windows,
// warp roles, queues and fragment rings) are described in banded_kernel.cuh and DESIGN.md section 4.
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include "magi_common.cuh"
#include "ode_models.cuh"

namespace magi {

// ------------------------------------------------------------------------------------------------------------
// K6: fragment tables.  fragtab[view][d][pair p][hh][lane][tt]: pair p holds the tiles J = 2p + tt, tt = 0, 1, interleaved per
// lane so that one 16-byte load fetches chunk hh of BOTH tiles -- the two DMMAs it feeds belong to different accumulate
// chains.  lane = 4*gid + q holds the B-operand entry
//   T[in = 4*(2J - HB + hh) + q][out = 8J + (gid>>1) + 4*(gid&1)]
// with T[in][out] = A[out][in] for y = A x (views 0: m~, 1: C~, 2: K~) and T[in][out] = m~[in][out] for view 3 (m~^T).
// Band rule |in - out| <= b as mat2band (gaussian_process.jl:70-74, 358-360).  Input tables are diagonal-major.
// ------------------------------------------------------------------------------------------------------------
size_t fragtab_doubles(int n, int b, int D) {
    BandGeom g = band_geom(n, b);
    return (size_t)4 * D * ((g.NT + 1) / 2) * g.NCH * 64;
}

__global__ void build_fragtab_kernel(const double* __restrict__ band_cinv, const double* __restrict__ band_mphi,
                                     const double* __restrict__ band_kinv, double* __restrict__ fragtab,
                                     int n, int b, int D, int HB, int NCH, int NT, int natural, double scale_c, double scale_k) {
    const int NP = (NT + 1) / 2;
    const size_t per_view = (size_t)D * NP * NCH * 64;
    const size_t total = 4 * per_view;
    const size_t tab = (size_t)(2 * b + 1) * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int within = (int)(idx % ((size_t)NCH * 64));       // position inside the (view, d, pair) block
        size_t r = idx / ((size_t)NCH * 64);
        const int tt = within & 1, lane = (within >> 1) & 31, hh = within >> 6;
        int p = (int)(r % NP); r /= NP;
        int d = (int)(r % D);
        int view = (int)(r / D);
        const int J = 2 * p + tt;
        int gid = lane >> 2, q = lane & 3;
        // windowed kernel: output slots permuted so that a C fragment is the next product's A fragment; dataflow kernel
        // (flow_kernel.cuh): natural order, B[k = q][n = gid] = T[in = 4c + q][out = 8J + gid]
        int o = natural ? 8 * J + gid : 8 * J + (gid >> 1) + 4 * (gid & 1);
        int i = 4 * (2 * J - HB + hh) + q;
        double v = 0.0;
        if (J >= 0 && J < NT && o < n && i >= 0 && i < n && abs(i - o) <= b) {
            const double* src = (view == 1) ? band_cinv : (view == 2 ? band_kinv : band_mphi);
            src += (size_t)d * tab;
            v = (view == 3) ? src[(size_t)(b + (o - i)) * n + i] : src[(size_t)(b + (i - o)) * n + o];
            if (view == 1) v *= scale_c;        // 1/beta2 folded into C~, 1/beta1 into K~: the kernels never multiply by them
            if (view == 2) v *= scale_k;
        }
        fragtab[idx] = v;
    }
}

cudaError_t launch_build_fragtab(const double* band_cinv, const double* band_mphi, const double* band_kinv, double* fragtab,
                                 int n, int b, int D, bool natural, double scale_c, double scale_k, cudaStream_t st) {
    BandGeom g = band_geom(n, b);
    size_t total = fragtab_doubles(n, b, D);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    build_fragtab_kernel<<<blocks, 256, 0, st>>>(band_cinv, band_mphi, band_kinv, fragtab, n, b, D, g.HB, g.NCH, g.NT, natural ? 1 : 0, scale_c, scale_k);
    return cudaGetLastError();
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

// ---- dataflow K1 (flow_kernel.cuh): shared-memory footprint, wavefront order ----
bool model_kx(int model, int& KX) {
    switch (model) {
    case MAGI_MODEL_FN: KX = Ode<MAGI_MODEL_FN>::KX; return true;
    case MAGI_MODEL_HES1: KX = Ode<MAGI_MODEL_HES1>::KX; return true;
    case MAGI_MODEL_HES1LOG: KX = Ode<MAGI_MODEL_HES1LOG>::KX; return true;
    case MAGI_MODEL_HES1LOG_FIXG: KX = Ode<MAGI_MODEL_HES1LOG_FIXG>::KX; return true;
    case MAGI_MODEL_HES1LOG_FIXF: KX = Ode<MAGI_MODEL_HES1LOG_FIXF>::KX; return true;
    case MAGI_MODEL_HIV: KX = Ode<MAGI_MODEL_HIV>::KX; return true;
    case MAGI_MODEL_PTRANS: KX = Ode<MAGI_MODEL_PTRANS>::KX; return true;
    case MAGI_MODEL_LV: KX = Ode<MAGI_MODEL_LV>::KX; return true;
    default: return false;
    }
}

// layout of flow_logpost_kernel's shared memory (G chain groups = 8 G chains per block)
size_t flow_smem_bytes(int D, int K, int KX, int n, int HB, int G, int& RS0) {
    const int NT = (n + 7) / 8, NP = (NT + 1) / 2, CH = 8 * G, RED = 3 + K;
    RS0 = (16 * NP + 8 * HB + 15) / 16 * 16;
    const size_t doubles = (size_t)3 * CH * D * RS0 + (size_t)D * NP * CH * RED + (size_t)2 * CH * (KX + 3 * D) + (size_t)CH * D * (RED + 4) + (size_t)D * 16 * NP;
    if (NP > 32) return (size_t)1 << 40;          // the final stage gives one lane to every tile pair
    return doubles * sizeof(double) + (size_t)2 * D * NP * sizeof(unsigned long long) + 16 + (size_t)3 * D * NP * sizeof(int);
}

// Order in which the warps draw the units (sweep s, dimension d, pair p).  Every unit a unit waits for comes strictly
// earlier in the list (S2 reads E of pairs p - HP .. p + HP, S3 reads KE of the same range and of the other dimensions), so
// drawing tickets in this order cannot deadlock.  extra_lag < 0: sweep after sweep (all S1, all S2, all S3) -- with 16 warps
// in flight a unit's producers are then long finished when its ticket is drawn.  extra_lag >= 0: a wavefront along the time
// axis (S1 of pair tau, S2 of pair tau - lag, S3 of pair tau - 2 lag with lag = HP + 1 + extra_lag), measured slower:
// producers and consumers a few tickets apart run concurrently and the consumers wait.
std::vector<int> flow_unit_order(int D, int NP, int HB, int extra_lag) {
    std::vector<int> u;
    if (extra_lag < 0) {
        for (int s = 0; s < 3; ++s)
            for (int p = 0; p < NP; ++p)
                for (int d = 0; d < D; ++d) u.push_back(s | (d << 2) | (p << 8));
        return u;
    }
    const int HP = (4 * HB + 15) / 16, lag = HP + 1 + extra_lag;
    for (int tau = 0; tau < NP + 2 * lag; ++tau)
        for (int s = 2; s >= 0; --s) {                      // the oldest (and heaviest) work of a wavefront step first
            const int p = tau - s * lag;
            if (p < 0 || p >= NP) continue;
            for (int d = 0; d < D; ++d) u.push_back(s | (d << 2) | (p << 8));
        }
    return u;
}

size_t banded_scratch_doubles_per_cta(int G, int D, int NT) { return (size_t)G * D * NT * 64; }   // Ke

// Chooses chain-groups per block (G <= gmax, a power of two) and where the Ke scratch lives (H and DW are fixed: one time
// segment, all dimensions concurrently).  Preference: 16 warps per block (one block per SM, <= 128 registers), as many
// chain-groups as possible sharing the fragment rings.
void banded_pick_config(int D, int K, int NT, int HB, int smem_limit, int gmax, int& G, int& H, int& DW, int& scratch_in_smem, size_t& smem_bytes) {
    const int NCH = 2 * HB + 2;
    (void)K;
    DW = D;
    H = 1;
    if (gmax < 1 || gmax > 4) gmax = 4;
    if (const char* e = getenv("MAGI_FORCE_G")) gmax = atoi(e);
    const bool force_global = getenv("MAGI_FORCE_GLOBAL_SCRATCH") != nullptr;
    // rings [D][kRingStages][4 blocks] + queues [G*D tasks][kXStages][kXSlots][32 lanes] + mbarriers (banded_kernel.cuh);
    // the per-chain reduction area aliases the queues
    constexpr int R = 3, S = 3, XS = 4;
    auto fixed_bytes = [&](int g) { return ((size_t)D * R * 4 * NCH * 32 + (size_t)g * D * S * XS * 32 + 4 * R * D + 4 * S * g * D) * sizeof(double); };
    // As many chain-groups per block as fit 16 warps (2 warps per task); for long time axes the Ke scratch of that many
    // groups does not fit shared memory and goes to global memory (L2-resident): on LV n=1281 G=4 with L2 scratch ran in
    // 0.27 ms against 0.45 ms for G=1 with shared-memory scratch.
    for (int g = gmax; g >= 1; --g) {
        if (2 * g * D > 16 || (g & (g - 1)) != 0) continue;      // powers of two only (the ring producer rotates with r & (G - 1))
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 0 && force_global) continue;
            size_t scr = pass == 0 ? banded_scratch_doubles_per_cta(g, D, NT) * sizeof(double) : 0;
            size_t tot = scr + fixed_bytes(g);      // (the per-chain reduction area aliases the queues)
            if (tot <= (size_t)smem_limit) { G = g; scratch_in_smem = (pass == 0); smem_bytes = tot; return; }
        }
    }
    G = 1; scratch_in_smem = 0;
    smem_bytes = fixed_bytes(1);
}

// one translation unit per model instantiates the kernels (banded_inst_*.cu), so they compile in parallel
#define MAGI_DECL_MODEL(M) cudaError_t launch_banded_model_##M(const BandedArgs& a, int HB, int DW, size_t smem_bytes, cudaStream_t st); \
                           cudaError_t launch_flow_model_##M(const FlowArgs& a, int HB, int grid, size_t smem_bytes, cudaStream_t st);
MAGI_DECL_MODEL(0) MAGI_DECL_MODEL(1) MAGI_DECL_MODEL(2) MAGI_DECL_MODEL(3) MAGI_DECL_MODEL(4) MAGI_DECL_MODEL(5) MAGI_DECL_MODEL(6) MAGI_DECL_MODEL(7)
#undef MAGI_DECL_MODEL
// `a.G`, `a.scratch_in_smem` must come from banded_pick_config; DW = blockDim warps / G.
cudaError_t launch_banded_cfg(int model, const BandedArgs& a, int HB, int DW, size_t smem_bytes, cudaStream_t st) {
    switch (model) {
    case MAGI_MODEL_FN: return launch_banded_model_0(a, HB, DW, smem_bytes, st);
    case MAGI_MODEL_HES1: return launch_banded_model_1(a, HB, DW, smem_bytes, st);
    case MAGI_MODEL_LV: return launch_banded_model_7(a, HB, DW, smem_bytes, st);
#ifndef MAGI_FAST_BUILD
    case MAGI_MODEL_HES1LOG: return launch_banded_model_2(a, HB, DW, smem_bytes, st);
    case MAGI_MODEL_HES1LOG_FIXG: return launch_banded_model_3(a, HB, DW, smem_bytes, st);
    case MAGI_MODEL_HES1LOG_FIXF: return launch_banded_model_4(a, HB, DW, smem_bytes, st);
    case MAGI_MODEL_HIV: return launch_banded_model_5(a, HB, DW, smem_bytes, st);
    case MAGI_MODEL_PTRANS: return launch_banded_model_6(a, HB, DW, smem_bytes, st);
#endif
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_flow_cfg(int model, const FlowArgs& a, int HB, int grid, size_t smem_bytes, cudaStream_t st) {
    switch (model) {
    case MAGI_MODEL_FN: return launch_flow_model_0(a, HB, grid, smem_bytes, st);
    case MAGI_MODEL_HES1: return launch_flow_model_1(a, HB, grid, smem_bytes, st);
    case MAGI_MODEL_LV: return launch_flow_model_7(a, HB, grid, smem_bytes, st);
#ifndef MAGI_FAST_BUILD
    case MAGI_MODEL_HES1LOG: return launch_flow_model_2(a, HB, grid, smem_bytes, st);
    case MAGI_MODEL_HES1LOG_FIXG: return launch_flow_model_3(a, HB, grid, smem_bytes, st);
    case MAGI_MODEL_HES1LOG_FIXF: return launch_flow_model_4(a, HB, grid, smem_bytes, st);
    case MAGI_MODEL_HIV: return launch_flow_model_5(a, HB, grid, smem_bytes, st);
    case MAGI_MODEL_PTRANS: return launch_flow_model_6(a, HB, grid, smem_bytes, st);
#endif
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace magi
