// Host side of K1-narrow (narrow_kernel.cuh): the per-step coefficient table and the model dispatch.
#include "magi_internal.cuh"
#include "narrow_kernel.cuh"

namespace magi {

#define NCK(call, what) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_error(e__, what); } while (0)

// two-component models (the windows of one chain fit a thread's registers up to b = 4)
bool narrow_supported(int model, int b) {
    return b >= 0 && b <= 4 && (model == MAGI_MODEL_FN || model == MAGI_MODEL_LV);
}

// the per-step coefficient table: built on first use and whenever the band tables (or the temperatures folded into it) change
int refresh_steptab(magi_handle* h, cudaStream_t st) {
    const int n = h->n, D = h->D, b = h->b;
    const int CS = narrow_cs(D, b), n_steps_pad = narrow_steps_pad(n, b);
    if (!h->d_steptab) { NCK(cudaMalloc(&h->d_steptab, sizeof(double) * (size_t)n_steps_pad * CS), "cudaMalloc step table"); h->steptab_dirty = true; }
    if (!h->steptab_dirty) return MAGI_OK;
    const size_t total = (size_t)n_steps_pad * CS;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    build_steptab_kernel<<<blocks, 256, 0, st>>>(h->d_band[0], h->d_band[1], h->d_band[2], h->d_yobs, h->d_steptab, n, b, D, CS, n_steps_pad,
                                                  1.0 / h->beta[1], 1.0 / h->beta[0]);
    NCK(cudaGetLastError(), "build_steptab_kernel");
    h->launches++; h->steptab_dirty = false;
    return MAGI_OK;
}

int eval_narrow_dev(magi_handle* h, int n_chains, const double* params, long long pitch, double* ll, double* grad, cudaStream_t st) {
    const int n = h->n, D = h->D, b = h->b;
    NarrowArgs a;
    a.CS = narrow_cs(D, b);
    a.n_steps_pad = narrow_steps_pad(n, b);
    int rc = refresh_steptab(h, st);
    if (rc) return rc;
    a.n = n; a.P = h->P; a.n_chains = n_chains; a.sigma_is_fixed = h->sigma_is_fixed; a.sigma_invalid = h->sigma_invalid;
    a.pitch = pitch; a.params = params; a.ll = ll; a.grad = grad; a.steptab = h->d_steptab;
    a.nobs = h->d_nobs; a.sigma_init = h->d_sigma_init; a.beta3 = h->beta[2]; a.inv_b3 = 1.0 / h->beta[2];
    cudaError_t e;
    switch (h->model) {
    case MAGI_MODEL_FN: e = narrow_launch_model<MAGI_MODEL_FN>(a, b, h->sm_count, st); break;
    case MAGI_MODEL_LV: e = narrow_launch_model<MAGI_MODEL_LV>(a, b, h->sm_count, st); break;
    default: e = cudaErrorInvalidValue;
    }
    NCK(e, "narrow_logpost_kernel launch");
    h->launches++;
    return MAGI_OK;
}

}  // namespace magi
