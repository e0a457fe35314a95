// Host side of K1-split (split_kernel.cuh): work space, slicing of large batches, model dispatch.
#include <cmath>
#include <cstdlib>
#include "magi_internal.cuh"
#include "split_kernel.cuh"

namespace magi {

#define SCK(call, what) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_error(e__, what); } while (0)

#define MAGI_DECL_SPLIT(M) cudaError_t launch_split_model_##M(const SplitArgs& a, int HB, int sm_count, cudaStream_t st, long long* launches);
MAGI_DECL_SPLIT(0) MAGI_DECL_SPLIT(1) MAGI_DECL_SPLIT(2) MAGI_DECL_SPLIT(3) MAGI_DECL_SPLIT(4) MAGI_DECL_SPLIT(5) MAGI_DECL_SPLIT(6) MAGI_DECL_SPLIT(7) MAGI_DECL_SPLIT(8)
#undef MAGI_DECL_SPLIT

static cudaError_t launch_split(int model, const SplitArgs& a, int HB, int sm_count, cudaStream_t st, long long* launches) {
    switch (model) {
    case MAGI_MODEL_FN: return launch_split_model_0(a, HB, sm_count, st, launches);
    case MAGI_MODEL_HES1: return launch_split_model_1(a, HB, sm_count, st, launches);
    case MAGI_MODEL_LV: return launch_split_model_7(a, HB, sm_count, st, launches);
#ifndef MAGI_FAST_BUILD
    case MAGI_MODEL_HES1LOG: return launch_split_model_2(a, HB, sm_count, st, launches);
    case MAGI_MODEL_HES1LOG_FIXG: return launch_split_model_3(a, HB, sm_count, st, launches);
    case MAGI_MODEL_HES1LOG_FIXF: return launch_split_model_4(a, HB, sm_count, st, launches);
    case MAGI_MODEL_HIV: return launch_split_model_5(a, HB, sm_count, st, launches);
    case MAGI_MODEL_PTRANS: return launch_split_model_6(a, HB, sm_count, st, launches);
    case MAGI_MODEL_L96: return launch_split_model_8(a, HB, sm_count, st, launches);
#endif
    default: return cudaErrorInvalidValue;
    }
}

// Chains per slice: the four planes of a slice should stay L2-resident between the stages (126 MB L2), but a slice must still
// give every SM work.
int split_slice_chains(const magi_handle* h) {
    const int HB = h->geom.HB, NP = (h->geom.NT + 1) / 2, TR = 2 * NP + 2 * ((HB + 1) / 2);
    const double per_chain = 4.0 * h->D * TR * 8 * sizeof(double);
    long long s = (long long)(72.0e6 / per_chain);
    s = s / 1024 * 1024;
    if (s < 1024) s = 1024;
    if (const char* e = getenv("MAGI_SPLIT_SLICE")) { const long long v = atoll(e); if (v >= 8) s = v / 8 * 8; }
    return (int)s;
}

int eval_split_dev(magi_handle* h, int n_chains, const double* params, long long pitch, double* ll, double* grad, cudaStream_t st) {
    const int n = h->n, D = h->D, HB = h->geom.HB;
    int KX = 0;
    if (h->model == MAGI_MODEL_L96) KX = 1; else model_kx(h->model, KX);
    // fragment tables with permuted output slots and 1/beta folded in (shared with the windowed kernel)
    if (!h->d_fragtab) { SCK(cudaMalloc(&h->d_fragtab, sizeof(double) * fragtab_doubles(n, h->b, D)), "cudaMalloc fragment tables"); h->frag_dirty = true; }
    if (h->frag_dirty) {
        SCK(launch_build_fragtab(h->d_band[0], h->d_band[1], h->d_band[2], h->d_fragtab, n, h->b, D, false, 1.0 / h->beta[1], 1.0 / h->beta[0], st), "build_fragtab");
        h->launches++; h->frag_dirty = false;
    }
    SplitArgs a;
    a.n = n; a.D = D; a.K = h->K; a.P = h->P; a.NP = (h->geom.NT + 1) / 2; a.MT = (HB + 1) / 2; a.TR = 2 * a.NP + 2 * a.MT; a.CW = KX + D;
    a.sigma_is_fixed = h->sigma_is_fixed; a.sigma_invalid = h->sigma_invalid;
    const int NV = 4 + h->K;
    int slice = split_slice_chains(h);
    if (slice > n_chains) slice = (n_chains + 7) / 8 * 8;
    if (slice > h->split_slice) {      // the layout (plane stride) follows the largest slice seen: the zero margin tiles are written once
        if (h->d_split_work) cudaFree(h->d_split_work);
        h->d_split_work = nullptr; h->split_slice = 0;
        const size_t need = 4 * (size_t)(slice / 8) * a.TR * 64 * D + (size_t)slice * D * a.NP * NV + (size_t)slice * a.CW;
        SCK(cudaMalloc(&h->d_split_work, sizeof(double) * need), "cudaMalloc K1-split work space");
        SCK(cudaMemsetAsync(h->d_split_work, 0, sizeof(double) * need, st), "memset K1-split work space");
        h->split_slice = slice;
    }
    const int cap = h->split_slice;
    a.NGc = cap / 8;
    a.plane = (long long)a.NGc * a.TR * 64;
    const size_t plane_all = (size_t)a.plane * D;
    a.XF = h->d_split_work;
    a.EF = a.XF + plane_all;
    a.KEF = a.EF + plane_all;
    a.G0F = a.KEF + plane_all;
    a.part = a.G0F + plane_all;
    a.cst = a.part + (size_t)cap * D * a.NP * NV;
    a.pitch = pitch; a.frag = h->d_fragtab;
    a.yobs = h->d_yobs; a.nobs = h->d_nobs; a.sigma_init = h->d_sigma_init;
    a.beta3 = h->beta[2]; a.inv_b3 = 1.0 / h->beta[2];
    for (int c0 = 0; c0 < n_chains; c0 += slice) {
        const int nc = (n_chains - c0 < slice) ? n_chains - c0 : slice;
        a.n_chains = nc; a.NG = (nc + 7) / 8; a.R = 1;
        a.params = params + (long long)c0 * pitch;
        a.ll = ll + c0;
        a.grad = grad ? grad + (long long)c0 * pitch : nullptr;
        SCK(launch_split(h->model, a, HB, h->sm_count, st, &h->launches), "K1-split launch");
    }
    return MAGI_OK;
}

}  // namespace magi
