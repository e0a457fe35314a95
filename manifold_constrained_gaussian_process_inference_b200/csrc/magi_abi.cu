// C ABI of libmagi_b200.so (see include/magi_b200.h for the reference interfaces each entry point replaces).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <limits>
#include <new>
#include "magi_common.cuh"
#include "magi_internal.cuh"

using namespace magi;

static thread_local std::string g_last_error;

namespace magi {
int set_error(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
int cuda_error(cudaError_t e, const char* what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return MAGI_ERR_CUDA;
}
}  // namespace magi

#define CK(call, what) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_error(e__, what); } while (0)

extern "C" const char* magi_last_error(void) { return g_last_error.c_str(); }
extern "C" int magi_version(void) { return 100; }

static void free_dev(void* p) { if (p) cudaFree(p); }

extern "C" int magi_destroy(magi_handle* h) {
    if (!h) return MAGI_OK;
    cudaSetDevice(h->device);
    for (int i = 0; i < 3; ++i) free_dev(h->d_band[i]);
    for (int i = 0; i < 7; ++i) free_dev(h->d_dense[i]);
    free_dev(h->d_fragtab); free_dev(h->d_fragtab_nat); free_dev(h->d_fragtab_bp); free_dev(h->d_steptab); free_dev(h->d_yobs); free_dev(h->d_nobs); free_dev(h->d_sigma_init);
    free_dev(h->d_params); free_dev(h->d_ll); free_dev(h->d_grad); free_dev(h->d_scratch);
    free_dev(h->d_dense_work); free_dev(h->d_dense_part); free_dev(h->d_dense_ops); free_dev(h->d_sk_work); free_dev(h->d_sk_flags);
    free_dev(h->d_small); free_dev(h->d_flow_units); if (h->h_pin) cudaFreeHost(h->h_pin);
    hmc_free(h);
    comm_free(h);
    if (h->stream) cudaStreamDestroy(h->stream);
    for (int i = 0; i < 3; ++i) if (h->pipe_streams[i]) cudaStreamDestroy(h->pipe_streams[i]);
    delete h;
    return MAGI_OK;
}

extern "C" int magi_create(const magi_config* cfg, magi_handle** out) {
    if (!cfg || !out) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_create: null argument");
    *out = nullptr;
    if (cfg->n_times < 1 || cfg->n_dims < 1) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_create: n_times and n_dims must be >= 1");
    if (!cfg->tvec || !cfg->yobs || !cfg->sigma_init || !cfg->prior_temperature)
        return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_create: tvec, yobs, sigma_init and prior_temperature are required");
    int mD = 0, mK = 0;
    if (cfg->ode_model_id == MAGI_MODEL_L96) { mD = cfg->n_dims; mK = 1; }
    else if (!model_dims(cfg->ode_model_id, mD, mK)) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_create: unknown ode_model_id");
    if (mD != cfg->n_dims || mK != cfg->n_params_ode) {
        char buf[160];
        snprintf(buf, sizeof buf, "magi_create: model %d has D=%d, k=%d but config says D=%d, k=%d", cfg->ode_model_id, mD, mK, cfg->n_dims, cfg->n_params_ode);
        return set_error(MAGI_ERR_INVALID_ARGUMENT, buf);
    }
    if (cfg->setup_mode != MAGI_SETUP_INJECT) {
        if (!cfg->phi) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_create: phi is required unless setup_mode is MAGI_SETUP_INJECT");
        if (cfg->kernel_id < MAGI_KERNEL_MATERN52 || cfg->kernel_id > MAGI_KERNEL_MATERN_NU52)
            return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_create: unknown kernel_id");
        if (cfg->kernel_id > MAGI_KERNEL_RBF)        // gaussian_process.jl:278-280: a warning, then the zero-derivative fallback
            g_last_error = "Time derivative calculation not implemented for this base kernel type. Derivatives will be zero.";
        for (int d = 0; d < cfg->n_dims; ++d) {   // src/MagiJl.jl:469-472
            double var = cfg->phi[2 * d], len = cfg->phi[2 * d + 1];
            if (!std::isfinite(var) || var <= 0 || !std::isfinite(len) || len <= 0)
                return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_create: invalid GP hyperparameters (variance and lengthscale must be finite and > 0)");
        }
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return set_error(MAGI_ERR_CUDA, "magi_create: no CUDA device available (libmagi_b200 has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_create: bad device ordinal");
    CK(cudaSetDevice(cfg->device), "cudaSetDevice");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device), "cudaGetDeviceProperties");
    if (prop.major != 10) return set_error(MAGI_ERR_UNSUPPORTED, "magi_create: libmagi_b200 is built for sm_100a (B200) only");

    magi_handle* h = new (std::nothrow) magi_handle();
    if (!h) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_create: out of host memory");
    h->n = cfg->n_times; h->D = cfg->n_dims; h->K = cfg->n_params_ode;
    h->sigma_is_fixed = cfg->sigma_is_fixed ? 1 : 0;
    h->P = h->n * h->D + h->K + (h->sigma_is_fixed ? 0 : h->D);
    int b = cfg->bandsize;
    if (b > h->n - 1) b = h->n - 1;             // src/MagiJl.jl:459-460
    if (b < 0) b = 0;
    h->b = b;
    h->kernel_id = cfg->kernel_id; h->model = cfg->ode_model_id; h->setup_mode = cfg->setup_mode;
    h->device = cfg->device; h->jitter = cfg->jitter;
    h->tvec.assign(cfg->tvec, cfg->tvec + h->n);
    if (cfg->phi) h->phi.assign(cfg->phi, cfg->phi + 2 * h->D);
    h->yobs.assign(cfg->yobs, cfg->yobs + (size_t)h->n * h->D);
    h->sigma_init.assign(cfg->sigma_init, cfg->sigma_init + h->D);
    for (int i = 0; i < 3; ++i) h->beta[i] = cfg->prior_temperature[i];
    h->sigma_invalid = 0;
    if (h->sigma_is_fixed)
        for (int d = 0; d < h->D; ++d)
            if (!std::isfinite(h->sigma_init[d]) || h->sigma_init[d] <= 0) h->sigma_invalid = 1;   // interface.jl:192
    h->geom = band_geom(h->n, h->b);
    h->smem_limit = (int)prop.sharedMemPerBlockOptin - 1024;
    h->sm_count = prop.multiProcessorCount;
    h->repaired_c.assign(h->D, 0); h->repaired_k.assign(h->D, 0);
    h->band_set.assign((size_t)3 * h->D, 0);

    int rc = MAGI_OK;
    auto fail = [&](int code) { magi_destroy(h); return code; };
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(set_error(MAGI_ERR_CUDA, "cudaStreamCreate failed"));
    const size_t tab = (size_t)(2 * h->b + 1) * h->n;
    for (int i = 0; i < 3; ++i) {
        if (cudaMalloc(&h->d_band[i], sizeof(double) * tab * h->D) != cudaSuccess) return fail(set_error(MAGI_ERR_CUDA, "cudaMalloc band tables failed"));
        cudaMemsetAsync(h->d_band[i], 0, sizeof(double) * tab * h->D, h->stream);
    }
    std::vector<int> nobs(h->D, 0);
    for (int d = 0; d < h->D; ++d)
        for (int i = 0; i < h->n; ++i) nobs[d] += std::isfinite(h->yobs[(size_t)d * h->n + i]) ? 1 : 0;
    if (cudaMalloc(&h->d_yobs, sizeof(double) * h->n * h->D) != cudaSuccess || cudaMalloc(&h->d_nobs, sizeof(int) * h->D) != cudaSuccess ||
        cudaMalloc(&h->d_sigma_init, sizeof(double) * h->D) != cudaSuccess)
        return fail(set_error(MAGI_ERR_CUDA, "cudaMalloc failed"));
    cudaMemcpyAsync(h->d_yobs, h->yobs.data(), sizeof(double) * h->n * h->D, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(h->d_nobs, nobs.data(), sizeof(int) * h->D, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(h->d_sigma_init, h->sigma_init.data(), sizeof(double) * h->D, cudaMemcpyHostToDevice, h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return fail(set_error(MAGI_ERR_CUDA, "upload failed"));

    h->dense_mode = (h->geom.HB > kMaxHB) || h->model == MAGI_MODEL_L96;
    if (!h->dense_mode) {
        banded_pick_config(h->D, h->K, h->geom.NT, h->geom.HB, h->smem_limit, 4, h->G, h->H, h->DW, h->scratch_in_smem, h->smem_bytes);
        // K1 variants: the windowed, warp-specialised kernel (banded_kernel.cuh) for large batches, the dataflow kernel
        // (flow_kernel.cuh: all 16 warps of a block work on 8 or 16 chains) for batches of at most one block per SM, where
        // the windowed kernel leaves most warps of the machine idle -- when X, E and KE of a block fit shared memory.
        // MAGI_K1=windowed|flow at create time forces one of them (development knob for A/B measurements).
        int KX = 0;
        model_kx(h->model, KX);
        for (int g = 1; g <= 2; ++g) {
            h->flow_smem[g] = flow_smem_bytes(h->D, h->K, KX, h->n, h->geom.HB, g, h->flow_RS0);
            h->flow_fits[g] = h->flow_smem[g] <= (size_t)prop.sharedMemPerBlockOptin;
        }
        const char* force = getenv("MAGI_K1");
        h->narrow_mode = (force && !strcmp(force, "narrow")) ? 1 : ((force && (!strcmp(force, "flow") || !strcmp(force, "windowed"))) ? -1 : 0);
        h->flow_mode = (force && !strcmp(force, "flow")) ? 1 : ((force && !strcmp(force, "windowed")) ? -1 : 0);
        if (h->flow_mode == 1 && !h->flow_fits[1]) return fail(set_error(MAGI_ERR_UNSUPPORTED, "MAGI_K1=flow: the state of 8 chains does not fit shared memory"));
        if (!h->flow_fits[1]) h->flow_mode = -1;
        if (h->flow_mode >= 0) {
            const char* ord = getenv("MAGI_FLOW_LAG");      // development knob: wavefront order with this extra lag
            std::vector<int> units = flow_unit_order(h->D, (h->geom.NT + 1) / 2, h->geom.HB, ord ? atoi(ord) : -1);
            h->flow_units = (int)units.size();
            if (cudaMalloc(&h->d_flow_units, sizeof(int) * units.size()) != cudaSuccess) return fail(set_error(MAGI_ERR_CUDA, "cudaMalloc unit order failed"));
            cudaMemcpy(h->d_flow_units, units.data(), sizeof(int) * units.size(), cudaMemcpyHostToDevice);
        }
    }
    if (h->setup_mode != MAGI_SETUP_INJECT) {
        rc = run_device_setup(h);
        if (rc != MAGI_OK) return fail(rc);
    }
    if (cfg->max_chains > 0) {
        rc = ensure_capacity(h, cfg->max_chains);
        if (rc != MAGI_OK) return fail(rc);
    }
    *out = h;
    return MAGI_OK;
}

namespace magi {

int ensure_capacity(magi_handle* h, int n_chains) {
    if ((size_t)n_chains <= h->cap_chains) return MAGI_OK;
    free_dev(h->d_params); free_dev(h->d_ll); free_dev(h->d_grad);
    h->d_params = h->d_ll = h->d_grad = nullptr; h->cap_chains = 0;
    size_t cap = (size_t)n_chains;
    CK(cudaMalloc(&h->d_params, sizeof(double) * cap * h->P), "cudaMalloc params staging");
    CK(cudaMalloc(&h->d_grad, sizeof(double) * cap * h->P), "cudaMalloc grad staging");
    CK(cudaMalloc(&h->d_ll, sizeof(double) * cap), "cudaMalloc ll staging");
    h->cap_chains = cap;
    return MAGI_OK;
}

// Block shape per call: four chain-groups per block share one fragment ring (best per-SM throughput), but a batch of at most
// 16 chains per SM would leave half the machine idle -- two groups per block then (tools/g_sweep.py: LV n=1281, 2048 chains
// 0.249 -> 0.191 ms; FN n=201, 2048 chains 0.057 -> 0.047 ms).
static void select_block_shape(magi_handle* h, int n_chains) {
    const int gmax = (n_chains <= 16 * h->sm_count) ? 2 : 4;
    if (gmax != h->gmax_cur) {
        banded_pick_config(h->D, h->K, h->geom.NT, h->geom.HB, h->smem_limit, gmax, h->G, h->H, h->DW, h->scratch_in_smem, h->smem_bytes);
        h->gmax_cur = gmax;
    }
}

static int ensure_scratch(magi_handle* h, int n_chains) {
    if (h->dense_mode || h->scratch_in_smem) return MAGI_OK;
    size_t blocks = ((size_t)n_chains + h->G * 8 - 1) / (h->G * 8);
    size_t need = blocks * banded_scratch_doubles_per_cta(h->G, h->D, h->geom.NT);
    if (need <= h->scratch_cap) return MAGI_OK;
    free_dev(h->d_scratch); h->d_scratch = nullptr; h->scratch_cap = 0;
    CK(cudaMalloc(&h->d_scratch, sizeof(double) * need), "cudaMalloc Ke scratch");
    h->scratch_cap = need;
    return MAGI_OK;
}

// Which K1 runs a batch of n_chains: 0 = windowed, else the dataflow kernel with that many chain groups per block.
static int flow_groups_for(const magi_handle* h, int n_chains_call) {
    if (h->flow_mode < 0) return 0;
    const long long n_chains = h->dispatch_chains > 0 ? h->dispatch_chains : n_chains_call;
    if (h->flow_mode > 0) return h->flow_fits[2] && n_chains > 8 * h->sm_count ? 2 : 1;
    if (n_chains <= 8 * h->sm_count) return 1;
    if (n_chains <= 16 * h->sm_count && h->flow_fits[2]) return 2;
    return 0;
}

// Band half-widths <= 4 of the two-component models: the FP64-FMA kernel (narrow_kernel.cuh) for batches that fill the machine.  One
// thread per chain needs the whole sweep whatever the batch (FN n=201: 43 / 49 / 84 us at b = 1 / 2 / 4), the windowed DMMA kernel 36 us
// per wave of 32 chains per SM: narrow wins where the windowed kernel needs a second wave (b <= 2) or a third one (b = 3, 4) --
// measured: b = 2, 8192 chains 0.049 against 0.072 ms; b = 4, 16 384 chains 0.088 against 0.104 ms; 4096 chains 0.049 / 0.078 against
// 0.036 ms.  Like the other variants it is chosen by the size of the CALL (or the sampler's global chain count), so how a batch is cut
// never changes the bits.
static bool narrow_route(const magi_handle* h, int n_chains_call) {
    if (h->narrow_mode < 0 || !narrow_supported(h->model, h->b)) return false;
    if (h->narrow_mode > 0) return true;
    const long long n_chains = h->dispatch_chains > 0 ? h->dispatch_chains : n_chains_call;
    return n_chains > (h->b <= 2 ? 32LL : 64LL) * h->sm_count;
}

// The two kernels read differently ordered fragment tables (the windowed kernel permutes the output slots of a tile); one
// buffer per layout, each built on first use and whenever the band tables change.
int refresh_fragtab(magi_handle* h, bool natural, cudaStream_t st) {
    if (h->dense_mode) return MAGI_OK;
    double*& tab = natural ? h->d_fragtab_nat : h->d_fragtab;
    bool& dirty = natural ? h->frag_nat_dirty : h->frag_dirty;
    if (!tab) { CK(cudaMalloc(&tab, sizeof(double) * fragtab_doubles(h->n, h->b, h->D)), "cudaMalloc fragment tables"); dirty = true; }
    if (!dirty) return MAGI_OK;
    CK(launch_build_fragtab(h->d_band[0], h->d_band[1], h->d_band[2], tab, h->n, h->b, h->D, natural,
                            1.0 / h->beta[1], 1.0 / h->beta[0], st), "build_fragtab");      // 1/beta2 folded into C~, 1/beta1 into K~
    h->launches++;
    dirty = false;
    return MAGI_OK;
}

// core evaluation on device pointers (chain-contiguous, pitch doubles per chain)
int eval_dev(magi_handle* h, int n_chains, const double* params_dev, long long pitch, double* ll_dev, double* grad_dev, cudaStream_t st) {
    if (n_chains <= 0) return MAGI_OK;
    if (!h->tables_ready) return set_error(MAGI_ERR_NOT_READY, "band tables are neither built nor fully injected (magi_set_band_tables for every dim and table)");
    if (h->dense_mode) return eval_dense_dev(h, n_chains, params_dev, pitch, ll_dev, grad_dev, st);
    if (narrow_route(h, n_chains)) return eval_narrow_dev(h, n_chains, params_dev, pitch, ll_dev, grad_dev, st);
    const int fg = flow_groups_for(h, n_chains);
    int rc = refresh_fragtab(h, fg > 0, st);
    if (rc) return rc;
    if (fg > 0) {
        FlowArgs f;
        f.G = fg;
        f.n = h->n; f.P = h->P; f.n_chains = n_chains; f.NP = (h->geom.NT + 1) / 2; f.RS0 = h->flow_RS0; f.n_units = h->flow_units;
        f.n_cblocks = (n_chains + 8 * fg - 1) / (8 * fg);
        f.sigma_is_fixed = h->sigma_is_fixed; f.sigma_invalid = h->sigma_invalid;
        f.pitch = pitch; f.params = params_dev; f.ll = ll_dev; f.grad = grad_dev;
        f.fragtab = h->d_fragtab_nat; f.units = h->d_flow_units; f.yobs = h->d_yobs; f.nobs = h->d_nobs; f.sigma_init = h->d_sigma_init;
        for (int i = 0; i < 3; ++i) { f.beta[i] = h->beta[i]; f.inv_beta[i] = 1.0 / h->beta[i]; }
        const int grid = f.n_cblocks < h->sm_count ? f.n_cblocks : h->sm_count;
        f.dbg = nullptr;
#ifdef MAGI_DEV_KNOBS      // development builds only (tools/build_variant.py): per-warp phase clocks of the dataflow kernel
        static const bool dbg_flow = getenv("MAGI_DBG_CLOCKS") != nullptr;
        long long* d_dbgf = nullptr;
        if (dbg_flow) { cudaMalloc(&d_dbgf, sizeof(long long) * 16 * 16 * grid); cudaMemset(d_dbgf, 0, sizeof(long long) * 16 * 16 * grid); f.dbg = d_dbgf; }
#endif
        CK(launch_flow_cfg(h->model, f, h->geom.HB, grid, h->flow_smem[fg], st), "flow_logpost_kernel launch");
        h->launches++;
#ifdef MAGI_DEV_KNOBS
        if (dbg_flow) {
            cudaStreamSynchronize(st);
            std::vector<long long> v((size_t)16 * 16 * grid);
            cudaMemcpy(v.data(), d_dbgf, sizeof(long long) * v.size(), cudaMemcpyDeviceToHost);
            double s[16] = {0};
            for (int i = 0; i < 16 * grid; ++i) for (int j = 0; j < 16; ++j) s[j] += (double)v[(size_t)i * 16 + j];
            const double nw = 16.0 * grid;
            fprintf(stderr, "[magi dbg flow] grid=%d chain blocks=%d  avg cycles per warp: load wait=%.0f loop=%.0f (of which dependency waits=%.0f) loop tail=%.0f final=%.0f units=%.1f\n",
                    grid, f.n_cblocks, s[0] / nw, s[1] / nw, s[2] / nw, s[3] / nw, s[4] / nw, s[5] / nw);
            fprintf(stderr, "[magi dbg flow] timeline (-DMAGI_FLOW_TIMELINE) arrive=%.0f S1: product=%.0f pointwise=%.0f | S2: wait=%.0f product=%.0f pointwise=%.0f | S3: product1=%.0f wait=%.0f product2=%.0f pointwise=%.0f\n",
                    s[6] / nw, s[7] / nw, s[8] / nw, s[9] / nw, s[10] / nw, s[11] / nw, s[12] / nw, s[13] / nw, s[14] / nw, s[15] / nw);
            cudaFree(d_dbgf);
        }
#endif
        return MAGI_OK;
    }
    select_block_shape(h, n_chains);
    rc = ensure_scratch(h, n_chains);
    if (rc) return rc;
    BandedArgs a;
    a.n = h->n; a.D = h->D; a.K = h->K; a.P = h->P; a.n_chains = n_chains; a.NT = h->geom.NT; a.G = h->G; a.H = h->H;
    a.sigma_is_fixed = h->sigma_is_fixed; a.sigma_invalid = h->sigma_invalid; a.scratch_in_smem = h->scratch_in_smem;
    a.pitch = pitch; a.params = params_dev; a.ll = ll_dev; a.grad = grad_dev;
    a.fragtab = h->d_fragtab; a.yobs = h->d_yobs; a.nobs = h->d_nobs; a.sigma_init = h->d_sigma_init;
    for (int i = 0; i < 3; ++i) { a.beta[i] = h->beta[i]; a.inv_beta[i] = 1.0 / h->beta[i]; }
    a.scratch = h->d_scratch;
    a.H = 0;
    a.dbg = nullptr;
#ifdef MAGI_DEV_KNOBS      // development builds only: A/B switch of the DMMA ping-pong, per-warp phase clocks
    { static const bool no_pp = getenv("MAGI_NO_PINGPONG") != nullptr; a.H = no_pp ? 1 : 0; }
    static const bool dbg_clocks = getenv("MAGI_DBG_CLOCKS") != nullptr;
    long long* d_dbg = nullptr;
    const int nblk = (n_chains + h->G * 8 - 1) / (h->G * 8), nwarp = h->G * h->DW;
    if (dbg_clocks) { cudaMalloc(&d_dbg, sizeof(long long) * 8 * nblk * nwarp); cudaMemset(d_dbg, 0, sizeof(long long) * 8 * nblk * nwarp); a.dbg = d_dbg; }
#endif
    CK(launch_banded_cfg(h->model, a, h->geom.HB, h->DW, h->smem_bytes, st), "banded_logpost_kernel launch");
    h->launches++;
#ifdef MAGI_DEV_KNOBS
    if (dbg_clocks) {
        cudaStreamSynchronize(st);
        std::vector<long long> v((size_t)8 * nblk * nwarp);
        cudaMemcpy(v.data(), d_dbg, sizeof(long long) * v.size(), cudaMemcpyDeviceToHost);
        double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < nblk * nwarp; ++i) { for (int j = 0; j < 8; ++j) s[j] += (double)v[i * 8 + j]; }
        const double nw = (double)nblk * nwarp;
        fprintf(stderr, "[magi dbg] blocks=%d tasks=%d G=%d smem=%zu  avg cycles per DMMA warp: A1=%.0f sync=%.0f A2=%.0f tail=%.0f (with -DMAGI_DBG_WAITS: DMMA warp waits ring=%.0f queue=%.0f, pointwise warp waits queue=%.0f)\n",
                nblk, nwarp, h->G, h->smem_bytes, s[0] / nw, s[1] / nw, s[2] / nw, s[3] / nw, s[4] / nw, s[5] / nw, s[7] / nw);
        cudaFree(d_dbg);
    }
#endif
    return MAGI_OK;
}

}  // namespace magi

extern "C" int magi_dimension(const magi_handle* h) { return h ? h->P : -1; }
extern "C" int magi_capabilities_order(const magi_handle* h) { (void)h; return 1; }
extern "C" long long magi_launch_count(const magi_handle* h) { return h ? h->launches : -1; }

extern "C" int magi_logdensity_and_gradient_batched_dev(magi_handle* h, int n_chains, const double* params_dev, double* ll_dev,
                                                        double* grad_dev, int layout, void* stream) {
    if (!h || !params_dev || !ll_dev) return set_error(MAGI_ERR_INVALID_ARGUMENT, "batched_dev: null argument");
    if (layout != MAGI_LAYOUT_CHAIN_CONTIGUOUS) return set_error(MAGI_ERR_UNSUPPORTED, "batched_dev: unknown layout");
    if (n_chains < 0) return set_error(MAGI_ERR_INVALID_ARGUMENT, "batched_dev: n_chains < 0");
    CK(cudaSetDevice(h->device), "cudaSetDevice");
    return eval_dev(h, n_chains, params_dev, h->P, ll_dev, grad_dev, (cudaStream_t)stream);
}

extern "C" int magi_logdensity_and_gradient_batched(magi_handle* h, int n_chains, const double* params, double* ll, double* grad) {
    if (!h || !params || !ll) return set_error(MAGI_ERR_INVALID_ARGUMENT, "batched: null argument");
    if (n_chains < 0) return set_error(MAGI_ERR_INVALID_ARGUMENT, "batched: n_chains < 0");
    if (n_chains == 0) return MAGI_OK;
    CK(cudaSetDevice(h->device), "cudaSetDevice");
    int rc = ensure_capacity(h, n_chains);
    if (rc) return rc;
    const size_t nb = sizeof(double) * (size_t)n_chains * h->P;
    // Large batches: chunks over three streams so that the H2D copy of chunk i+1, the kernel of chunk i and the D2H copy of
    // chunk i-1 overlap (PCIe is full duplex; with pinned host buffers the call is bound by one direction of the link
    // instead of H2D + kernel + D2H in sequence).  The banded kernel keeps its scratch in shared memory, so chunks are
    // independent; other configurations take the single-stream path.
    const int chunk_min = 1024;
    int nchunks = n_chains / chunk_min; if (nchunks > 8) nchunks = 8;
#ifdef MAGI_DEV_KNOBS
    { static const char* e = getenv("MAGI_E2E_CHUNKS"); if (e && atoi(e) > 0) nchunks = atoi(e); }
#endif
    const int per = nchunks > 0 ? ((n_chains + nchunks - 1) / nchunks + 31) / 32 * 32 : n_chains;
    // the K1 variant is chosen by the size of the CALL, not of its chunks: how a batch is cut into chunks never changes the result
    struct DispatchPin { magi_handle* h; long long old; ~DispatchPin() { h->dispatch_chains = old; } } pin{h, h->dispatch_chains};
    if (h->dispatch_chains == 0) h->dispatch_chains = n_chains;
    const bool chunk_flow = !h->dense_mode && flow_groups_for(h, per) > 0;
    if (!h->dense_mode && !chunk_flow && h->tables_ready) select_block_shape(h, per);      // the shape every chunk will run with
    if (n_chains >= 2 * chunk_min && !h->dense_mode && (h->scratch_in_smem || chunk_flow) && h->tables_ready) {
        for (int i = 0; i < 3; ++i)
            if (!h->pipe_streams[i]) CK(cudaStreamCreateWithFlags(&h->pipe_streams[i], cudaStreamNonBlocking), "cudaStreamCreate");
        rc = narrow_route(h, n_chains) ? refresh_steptab(h, h->stream) : refresh_fragtab(h, chunk_flow, h->stream);      // built before the chunks' streams use it
        if (rc) return rc;
        CK(cudaStreamSynchronize(h->stream), "stream sync");
        int i = 0;
        for (int c0 = 0; c0 < n_chains; c0 += per, ++i) {
            const int nc = (n_chains - c0 < per) ? n_chains - c0 : per;
            cudaStream_t st = h->pipe_streams[i % 3];
            const size_t off = (size_t)c0 * h->P;
            CK(cudaMemcpyAsync(h->d_params + off, params + off, sizeof(double) * (size_t)nc * h->P, cudaMemcpyHostToDevice, st), "H2D params");
            rc = eval_dev(h, nc, h->d_params + off, h->P, h->d_ll + c0, grad ? h->d_grad + off : nullptr, st);
            if (rc) return rc;
            CK(cudaMemcpyAsync(ll + c0, h->d_ll + c0, sizeof(double) * nc, cudaMemcpyDeviceToHost, st), "D2H ll");
            if (grad) CK(cudaMemcpyAsync(grad + off, h->d_grad + off, sizeof(double) * (size_t)nc * h->P, cudaMemcpyDeviceToHost, st), "D2H grad");
        }
        for (int k = 0; k < 3; ++k) CK(cudaStreamSynchronize(h->pipe_streams[k]), "stream sync");
        return MAGI_OK;
    }
    // Small calls (the single-chain drop-in of the reference's NUTS loop): pinned staging on both sides and ONE device-to-host
    // copy of a contiguous [ll | grad] block instead of two copies into pageable memory.
    const size_t np_d = (size_t)n_chains * h->P, nl_pad = ((size_t)n_chains + 7) / 8 * 8;
    if (np_d + nl_pad <= (size_t)65536) {
        const size_t half = 65536 + 8;
        if (!h->h_pin) {
            CK(cudaMallocHost(&h->h_pin, sizeof(double) * 2 * half), "cudaMallocHost staging");
            CK(cudaMalloc(&h->d_small, sizeof(double) * 2 * half), "cudaMalloc small-call staging");
            h->small_cap = half;
        }
        double* d_in = h->d_small;
        double* d_out = h->d_small + half;               // [ll (padded to 8) | grad]
        memcpy(h->h_pin, params, nb);
        CK(cudaMemcpyAsync(d_in, h->h_pin, nb, cudaMemcpyHostToDevice, h->stream), "H2D params");
        rc = eval_dev(h, n_chains, d_in, h->P, d_out, grad ? d_out + nl_pad : nullptr, h->stream);
        if (rc) return rc;
        const size_t out_d = grad ? nl_pad + np_d : (size_t)n_chains;
        CK(cudaMemcpyAsync(h->h_pin + half, d_out, sizeof(double) * out_d, cudaMemcpyDeviceToHost, h->stream), "D2H ll | grad");
        CK(cudaStreamSynchronize(h->stream), "stream sync");
        memcpy(ll, h->h_pin + half, sizeof(double) * n_chains);
        if (grad) memcpy(grad, h->h_pin + half + nl_pad, nb);
        return MAGI_OK;
    }
    CK(cudaMemcpyAsync(h->d_params, params, nb, cudaMemcpyHostToDevice, h->stream), "H2D params");
    rc = eval_dev(h, n_chains, h->d_params, h->P, h->d_ll, grad ? h->d_grad : nullptr, h->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ll, h->d_ll, sizeof(double) * n_chains, cudaMemcpyDeviceToHost, h->stream), "D2H ll");
    if (grad) CK(cudaMemcpyAsync(grad, h->d_grad, nb, cudaMemcpyDeviceToHost, h->stream), "D2H grad");
    CK(cudaStreamSynchronize(h->stream), "stream sync");
    return MAGI_OK;
}

extern "C" int magi_logdensity_and_gradient(magi_handle* h, const double* params, int n_params, double* ll, double* grad) {
    if (!h || !ll) return set_error(MAGI_ERR_INVALID_ARGUMENT, "logdensity_and_gradient: null argument");
    if (n_params != h->P || !params) {          // interface.jl:179-182
        *ll = -std::numeric_limits<double>::infinity();
        if (grad) for (int i = 0; i < h->P; ++i) grad[i] = std::numeric_limits<double>::quiet_NaN();
        g_last_error = "Dimension mismatch in logdensity_and_gradient";
        return MAGI_OK;
    }
    return magi_logdensity_and_gradient_batched(h, 1, params, ll, grad);
}

extern "C" int magi_logdensity(magi_handle* h, const double* params, int n_params, double* ll) {
    if (!h || !ll) return set_error(MAGI_ERR_INVALID_ARGUMENT, "logdensity: null argument");
    if (n_params != h->P || !params) {          // interface.jl:113-116
        *ll = -std::numeric_limits<double>::infinity();
        g_last_error = "Dimension mismatch in logdensity";
        return MAGI_OK;
    }
    // the reference computes the gradient and discards it (interface.jl:148); the kernel skips the stores
    return magi_logdensity_and_gradient_batched(h, 1, params, ll, nullptr);
}

extern "C" int magi_set_band_tables(magi_handle* h, int dim, int which, const double* in) {
    if (!h || !in) return set_error(MAGI_ERR_INVALID_ARGUMENT, "set_band_tables: null argument");
    if (dim < 0 || dim >= h->D) return set_error(MAGI_ERR_INVALID_ARGUMENT, "set_band_tables: dim out of range");
    if (which < MAGI_MAT_CINV_BAND || which > MAGI_MAT_KINV_BAND) return set_error(MAGI_ERR_INVALID_ARGUMENT, "set_band_tables: which must be a *_BAND selector");
    CK(cudaSetDevice(h->device), "cudaSetDevice");
    const size_t tab = (size_t)(2 * h->b + 1) * h->n;
    const int t = which - MAGI_MAT_CINV_BAND;
    CK(cudaMemcpy(h->d_band[t] + (size_t)dim * tab, in, sizeof(double) * tab, cudaMemcpyHostToDevice), "H2D band table");
    h->band_set[(size_t)t * h->D + dim] = 1;
    h->frag_dirty = true; h->frag_nat_dirty = true; h->frag_bp_dirty = true; h->steptab_dirty = true;
    h->dense_band_dirty = true;
    bool all = true;
    for (char c : h->band_set) all = all && c;
    if (all) h->tables_ready = true;
    return MAGI_OK;
}

extern "C" int magi_get_matrix(magi_handle* h, int dim, int which, double* out) {
    if (!h || !out) return set_error(MAGI_ERR_INVALID_ARGUMENT, "get_matrix: null argument");
    if (dim < 0 || dim >= h->D) return set_error(MAGI_ERR_INVALID_ARGUMENT, "get_matrix: dim out of range");
    CK(cudaSetDevice(h->device), "cudaSetDevice");
    if (which >= MAGI_MAT_CINV_BAND && which <= MAGI_MAT_KINV_BAND) {
        const size_t tab = (size_t)(2 * h->b + 1) * h->n;
        CK(cudaMemcpy(out, h->d_band[which - MAGI_MAT_CINV_BAND] + (size_t)dim * tab, sizeof(double) * tab, cudaMemcpyDeviceToHost), "D2H band table");
        return MAGI_OK;
    }
    if (which < 0 || which > MAGI_MAT_KINV) return set_error(MAGI_ERR_INVALID_ARGUMENT, "get_matrix: unknown selector");
    if (!h->d_dense[which]) return set_error(MAGI_ERR_NOT_READY, "get_matrix: dense matrices exist only after a device setup (setup_mode != INJECT)");
    const size_t nn = (size_t)h->n * h->n;
    CK(cudaMemcpy(out, h->d_dense[which] + (size_t)dim * nn, sizeof(double) * nn, cudaMemcpyDeviceToHost), "D2H dense matrix");
    return MAGI_OK;
}

extern "C" int magi_setup_timing(const magi_handle* h, double* kernel_ms, double* alloc_ms) {
    if (!h) return set_error(MAGI_ERR_INVALID_ARGUMENT, "setup_timing: null handle");
    if (kernel_ms) *kernel_ms = h->setup_kernel_ms;
    if (alloc_ms) *alloc_ms = h->setup_alloc_ms;
    return MAGI_OK;
}

extern "C" int magi_setup_status(magi_handle* h, int dim, int* rc_, int* rk_) {
    if (!h || dim < 0 || dim >= h->D) return set_error(MAGI_ERR_INVALID_ARGUMENT, "setup_status: bad argument");
    if (rc_) *rc_ = h->repaired_c[dim];
    if (rk_) *rk_ = h->repaired_k[dim];
    return MAGI_OK;
}
