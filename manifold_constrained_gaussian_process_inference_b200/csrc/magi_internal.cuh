// Internal handle layout and cross-file entry points of libmagi_b200.so.
#pragma once
#include "magi_common.cuh"

struct magi_handle {
    int n = 0, D = 0, K = 0, P = 0, b = 0;
    int kernel_id = 0, model = 0, sigma_is_fixed = 0, sigma_invalid = 0, setup_mode = 0, device = 0;
    double jitter = 1e-6;
    double beta[3] = {1.0, 1.0, 1.0};
    std::vector<double> tvec, phi, yobs, sigma_init;
    magi::BandGeom geom{};
    // device-resident GP tables.  d_band[0..2] = CinvBand, mphiBand, KinvBand, each D x (2b+1) x n diagonal-major
    double* d_band[3] = {nullptr, nullptr, nullptr};
    // dense GPCov fields (C, Cinv, Cprime, Cdoubleprime, mphi, Kphi, Kinv), each D x n x n column-major; only after a device setup
    double* d_dense[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double* d_fragtab = nullptr;       // windowed kernel's layout
    double* d_fragtab_nat = nullptr;   // dataflow kernel's (natural) layout
    double* d_fragtab_bp = nullptr;    // band-product route of many-component models: natural layout, 1/beta not folded in
    bool frag_bp_dirty = true;
    double* d_steptab = nullptr;       // K1-narrow (narrow_kernel.cuh): per-step coefficient table of band half-widths <= 4
    bool steptab_dirty = true;
    int narrow_mode = 0;               // 1: always when supported (MAGI_K1=narrow), -1: never, 0: by batch size
    double* d_yobs = nullptr;
    int* d_nobs = nullptr;
    double* d_sigma_init = nullptr;
    std::vector<char> band_set;
    bool tables_ready = false, frag_dirty = true, frag_nat_dirty = true, dense_band_dirty = true;
    bool dense_mode = false;
    // host-API staging
    double *d_params = nullptr, *d_ll = nullptr, *d_grad = nullptr;
    size_t cap_chains = 0;
    // small host-buffer calls (the single-chain drop-in): pinned staging and one contiguous [ll | grad] output block
    double *h_pin = nullptr, *d_small = nullptr;
    size_t small_cap = 0;                 // doubles per staging half
    double* d_scratch = nullptr;
    size_t scratch_cap = 0;
    // dense-mode work space
    double* d_dense_ops = nullptr;     // band-truncated dense Cinv~, mphi~, Kinv~ ([3][D][n x n]) for the dense path
    double* d_dense_work = nullptr;
    double* d_dense_part = nullptr;    // [n_chains][D][4 + K] per-(chain, dimension) partial sums of the dense route's gradient stage
    size_t dense_part_cap = 0;
    size_t dense_work_cap = 0;
    double* d_sk_work = nullptr;       // stream-K partial tiles and flags (gemm_f64.cuh)
    unsigned* d_sk_flags = nullptr;
    unsigned sk_epoch = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t pipe_streams[3] = {nullptr, nullptr, nullptr};   // H2D / kernel / D2H overlap in the host-buffer batched call
    long long launches = 0;
    int smem_limit = 0, sm_count = 148;
    int G = 2, H = 1, DW = 1, scratch_in_smem = 1, gmax_cur = 4;
    size_t smem_bytes = 0;
    // dataflow K1 (flow_kernel.cuh): chosen at create when the state of 16 chains fits shared memory
    long long dispatch_chains = 0;     // > 0: choose the K1 variant as for a batch of this many chains (the sampler sets the GLOBAL
                                       // chain count of a multi-rank run, so that every sharding runs the same kernel: bit-identical draws)
    int flow_mode = 0;                 // 0: by batch size (small batches), 1: always (MAGI_K1=flow), -1: never (MAGI_K1=windowed)
    bool flow_fits[3] = {false, false, false};      // [G]: the state of 8 G chains fits shared memory
    int flow_RS0 = 0, flow_units = 0;
    size_t flow_smem[3] = {0, 0, 0};
    int* d_flow_units = nullptr;
    std::vector<int> repaired_c, repaired_k;
    double setup_alloc_ms = 0.0, setup_kernel_ms = 0.0;   // device setup: host time in cudaMalloc / device time of K3-K6
    void* hmc = nullptr;     // on-device sampler state (hmc.cu)
    // NCCL communicator of a multi-rank run (comm.cu): created from a unique id (owned) or attached by the host
    void* nccl_comm = nullptr;
    bool nccl_owned = false;
    int nccl_rank = 0, nccl_world = 1;
};

namespace magi {
int set_error(int code, const std::string& msg);
int cuda_error(cudaError_t e, const char* what);
int ensure_capacity(magi_handle* h, int n_chains);
int refresh_fragtab(magi_handle* h, bool natural, cudaStream_t st);
int eval_dev(magi_handle* h, int n_chains, const double* params_dev, long long pitch, double* ll_dev, double* grad_dev, cudaStream_t st);
bool narrow_supported(int model, int b);
int refresh_steptab(magi_handle* h, cudaStream_t st);
int eval_narrow_dev(magi_handle* h, int n_chains, const double* params_dev, long long pitch, double* ll_dev, double* grad_dev, cudaStream_t st);
int eval_dense_dev(magi_handle* h, int n_chains, const double* params_dev, long long pitch, double* ll_dev, double* grad_dev, cudaStream_t st);
int run_device_setup(magi_handle* h);
void hmc_free(magi_handle* h);
void comm_free(magi_handle* h);
int comm_allreduce_sum(magi_handle* h, double* buf, size_t n, cudaStream_t st);
int comm_allgather(magi_handle* h, const double* send, double* recv, size_t n_per_rank, cudaStream_t st);
cudaError_t launch_banded_cfg(int model, const BandedArgs& a, int HB, int DW, size_t smem_bytes, cudaStream_t st);
cudaError_t launch_band_product(const double* fragtab, int view, const double* in, long long cs, long long ds, double* out, long long plane,
                                int n, int b, int D, int n_chains, int sm_count, cudaStream_t st);
cudaError_t launch_flow_cfg(int model, const FlowArgs& a, int HB, int grid, size_t smem_bytes, cudaStream_t st);
}  // namespace magi
