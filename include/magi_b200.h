/* magi_b200.h -- C ABI of libmagi_b200.so: the B200-native MAGI log-posterior + gradient path.
 *
 * This is the drop-in boundary for the reference's hot path (MagiJl.jl; citations are file:line under
 * the reference repository k1m9l/MAnifold_Constrained_Gaussian_Process_Inference):
 *
 *   magi_create                         replaces  GPCov() + calculate_gp_covariances!  src/gaussian_process.jl:14-54, 219-363
 *                                                 (called per dimension at src/MagiJl.jl:456-491)
 *                                       and       MagiTarget(...) construction         src/logdensityproblems_interface.jl:33-45, src/MagiJl.jl:508-520
 *   magi_dimension                      replaces  LogDensityProblems.dimension         src/logdensityproblems_interface.jl:53-61
 *   magi_capabilities_order             replaces  LogDensityProblems.capabilities      src/logdensityproblems_interface.jl:68-70
 *   magi_logdensity                     replaces  LogDensityProblems.logdensity        src/logdensityproblems_interface.jl:111-166
 *   magi_logdensity_and_gradient        replaces  LogDensityProblems.logdensity_and_gradient  src/logdensityproblems_interface.jl:176-267
 *                                                 (which wraps log_likelihood_and_gradient_banded, src/likelihoods.jl:43-257)
 *   magi_logdensity_and_gradient_batched[_dev]    the same evaluation for many independent chains (new: the reference
 *                                                 runs one chain, src/samplers.jl:173-184)
 *   magi_get_matrix / magi_set_band_tables        read / inject the GPCov fields         src/gaussian_process.jl:21-33
 *   magi_hmc_*                          on-device batched HMC transitions: the caller either side of the hot path
 *                                                 (run_nuts_sampler, src/samplers.jl:114-194), SURVEY.md section 8(f) row 1
 *
 * Conventions: every pointer is a HOST pointer unless the name ends in _dev; all floating point is IEEE
 * binary64; matrices are column-major (Julia layout); every function returns 0 on success and a non-zero
 * magi_status otherwise, with a message available from magi_last_error() (thread-local).  No exception and no
 * callback crosses this boundary.  The caller owns every buffer it passes; the library owns device memory until
 * magi_destroy.  One handle belongs to one (process, GPU); calls on one handle must be serialised by the caller,
 * different handles may be used concurrently.  There is no CPU fallback: without a CUDA device magi_create fails.
 */
#ifndef MAGI_B200_H
#define MAGI_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct magi_handle magi_handle;

typedef enum {
    MAGI_OK = 0,
    MAGI_ERR_INVALID_ARGUMENT = 1,
    MAGI_ERR_CUDA = 2,
    MAGI_ERR_NOT_READY = 3,       /* band tables neither built nor injected */
    MAGI_ERR_UNSUPPORTED = 4,
    MAGI_ERR_NOT_POSITIVE_DEFINITE = 5
} magi_status;

/* kernel ids: src/kernels.jl:74-81 (matern52), :42-50 (rbf), :109-118 (general Matern, nu = 1/2, 3/2, 5/2 in closed form).
 * The reference has analytic time derivatives for Matern52Kernel and SqExponentialKernel only; every other base kernel --
 * MaternKernel(nu) included, even at nu = 5/2 -- gets C but zero derivatives, i.e. the fallback m = 0, K = eI, Kinv = I/e
 * (src/gaussian_process.jl:278-280, 319-331).  The three MATERN_NU* ids reproduce that behaviour. */
enum { MAGI_KERNEL_MATERN52 = 0, MAGI_KERNEL_RBF = 1, MAGI_KERNEL_MATERN_NU12 = 2, MAGI_KERNEL_MATERN_NU32 = 3, MAGI_KERNEL_MATERN_NU52 = 4 };

/* compiled ODE model registry (user callbacks cannot cross a C ABI into a kernel); src/ode_models.jl */
enum {
    MAGI_MODEL_FN = 0,            /* fn_ode!            :39-47,  Jacobians :248-299 */
    MAGI_MODEL_HES1 = 1,          /* hes1_ode!          :60-70,  Jacobians :312-378 */
    MAGI_MODEL_HES1LOG = 2,       /* hes1log_ode!       :83-103 (Jacobians derived: the reference has none) */
    MAGI_MODEL_HES1LOG_FIXG = 3,  /* hes1log_ode_fixg!  :116-135 */
    MAGI_MODEL_HES1LOG_FIXF = 4,  /* hes1log_ode_fixf!  :147-165 */
    MAGI_MODEL_HIV = 5,           /* hiv_ode!           :178-207 */
    MAGI_MODEL_PTRANS = 6,        /* ptrans_ode!        :219-233 */
    MAGI_MODEL_LV = 7,            /* Lotka-Volterra (not in the reference; BASELINE config 3) */
    MAGI_MODEL_L96 = 8            /* Lorenz-96, D components (not in the reference; BASELINE config 4) */
};

/* how magi_create obtains Cinv / mphi / Kinv */
enum {
    MAGI_SETUP_REFERENCE_ORDER = 0,  /* Cinv = inv(chol(C+eI)); m = C'Cinv; K = C'' - m C'^T + eI; Kinv = inv(chol(K))  (gaussian_process.jl:295-318) */
    MAGI_SETUP_STABLE = 1,           /* W = L^-1 C'^T; K = C'' - W^T W + eI; m = (L^-T W)^T : same mathematics, PD by construction */
    MAGI_SETUP_INJECT = 2            /* no device setup: band tables must be injected with magi_set_band_tables */
};

/* matrix selectors for magi_get_matrix / magi_set_band_tables (GPCov fields, gaussian_process.jl:21-33) */
enum {
    MAGI_MAT_C = 0, MAGI_MAT_CINV = 1, MAGI_MAT_CPRIME = 2, MAGI_MAT_CDOUBLEPRIME = 3, MAGI_MAT_MPHI = 4,
    MAGI_MAT_KPHI = 5, MAGI_MAT_KINV = 6,                       /* dense n x n, column-major */
    MAGI_MAT_CINV_BAND = 7, MAGI_MAT_MPHI_BAND = 8, MAGI_MAT_KINV_BAND = 9   /* (2b+1) x n diagonal-major: T[(b + j - i) * n + i] = A[i][j] */
};

/* device layouts accepted by the *_dev entry points */
enum { MAGI_LAYOUT_CHAIN_CONTIGUOUS = 0 /* params[c * P + p]: a Julia P x n_chains Matrix */ };

typedef struct {
    int n_times;             /* n */
    int n_dims;              /* D */
    int n_params_ode;        /* k (checked against the model) */
    int kernel_id;           /* MAGI_KERNEL_* */
    int bandsize;            /* clamped to n-1 like src/MagiJl.jl:459; n-1 selects the dense path */
    int ode_model_id;        /* MAGI_MODEL_* */
    int sigma_is_fixed;      /* MagiTarget.sigma_is_fixed */
    int setup_mode;          /* MAGI_SETUP_* */
    int max_chains;          /* capacity hint for scratch (0 = grow on demand) */
    int device;              /* CUDA device ordinal */
    double jitter;           /* config[:jitter], default 1e-6 (src/MagiJl.jl:218) */
    const double* tvec;      /* n */
    const double* phi;       /* 2 x D column-major: phi[2*d] = variance, phi[2*d+1] = lengthscale (src/MagiJl.jl:466-467) */
    const double* yobs;      /* n x D column-major, non-finite = missing (src/likelihoods.jl:123) */
    const double* sigma_init;        /* D */
    const double* prior_temperature; /* 3: [deriv, level, obs] */
} magi_config;

const char* magi_last_error(void);
int magi_version(void);

int magi_create(const magi_config* cfg, magi_handle** out);
int magi_destroy(magi_handle* h);

int magi_dimension(const magi_handle* h);              /* P = nD + k (+ D if sigma is sampled); < 0 on error */
int magi_capabilities_order(const magi_handle* h);     /* 1  (LogDensityOrder{1}) */

/* single chain, exact reference semantics including the -Inf / zero-gradient guards.  n_params must equal
 * magi_dimension(h); otherwise *ll = -Inf and grad is filled with NaN (interface.jl:179-182), status MAGI_OK. */
int magi_logdensity(magi_handle* h, const double* params, int n_params, double* ll);
int magi_logdensity_and_gradient(magi_handle* h, const double* params, int n_params, double* ll, double* grad);

/* many chains; params is P x n_chains column-major (chain-contiguous); guards applied per chain.
 * grad may be NULL (value only). */
int magi_logdensity_and_gradient_batched(magi_handle* h, int n_chains, const double* params, double* ll, double* grad);

/* device-resident variant: no host copies, asynchronous on `stream` (a cudaStream_t passed as void*). */
int magi_logdensity_and_gradient_batched_dev(magi_handle* h, int n_chains, const double* params_dev, double* ll_dev,
                                             double* grad_dev, int layout, void* stream);

/* GPCov read-back / injection.  `out` / `in` hold n*n doubles (dense) or (2b+1)*n doubles (band). */
int magi_get_matrix(magi_handle* h, int dim, int which, double* out);
int magi_set_band_tables(magi_handle* h, int dim, int which, const double* in);
/* status of the device setup for one dimension: repaired (non-positive) pivots seen in chol(C+eI), chol(K+eI) */
int magi_setup_status(magi_handle* h, int dim, int* repaired_pivots_c, int* repaired_pivots_k);

/* where the time of the device setup inside magi_create went: device time from the covariance build to the band tables (K3-K6,
 * CUDA events on the handle's stream) and host time spent in cudaMalloc for the GPCov fields and the work space */
int magi_setup_timing(const magi_handle* h, double* kernel_ms, double* alloc_ms);

/* Stand-alone GPCov for ONE dimension, computed on the GPU: replaces calculate_gp_covariances!(gp_cov, kernel, phi, tvec,
 * bandsize; complexity, jitter) (src/gaussian_process.jl:219-363).  phi = [variance, lengthscale]; outputs (any may be
 * NULL): seven dense n x n column-major matrices and three (2b+1) x n band tables; repaired[2] = repaired pivot counts. */
int magi_gp_covariances(int kernel_id, const double* phi, const double* tvec, int n, int bandsize, double jitter,
                        int complexity, int setup_mode, int device, double* C, double* Cinv, double* Cprime,
                        double* Cdoubleprime, double* mphi, double* Kphi, double* Kinv, double* CinvBand,
                        double* mphiBand, double* KinvBand, int* repaired);

/* GP hyper-parameter initialisation objective: negative log marginal likelihood of
 * Initialization.negative_log_marginal_likelihood (src/initialization.jl:72-176) for n_cand candidate vectors
 * (log variance, log lengthscale, log sigma) on the n finite observations (t, y) of one dimension.  Invalid parameters or
 * non-finite results give +Inf, as in the reference. */
int magi_gp_nlml_batched(int kernel_id, int n, const double* t, const double* y, double jitter, int n_cand,
                         const double* log_params /* 3 x n_cand */, double* out /* n_cand */, int device);

/* ---- on-device batched HMC: the sampler loop of run_nuts_sampler (src/samplers.jl:114-194:
 * diagonal Euclidean metric, leapfrog, Stan-style step-size and metric adaptation) with the chain state resident in HBM.
 * Static trajectories of n_leapfrog steps; every leapfrog step is one evaluation of the hot path for every chain.
 * Random numbers are Philox streams keyed by (seed, chain_id_offset + chain), so a run does not depend on how the
 * chains are sharded over GPUs. */
int magi_hmc_init(magi_handle* h, int n_chains, const double* params0 /* P x n_chains */, unsigned long long seed,
                  double step_size0, long long chain_id_offset);
/* multi-rank runs (chains sharded over GPUs): total chain count and an in-place sum-over-ranks callback
 * (device pointer, number of doubles, cudaStream_t, user) for the pooled window statistics of the warm-up; returns 0 on success */
typedef int (*magi_allreduce_fn)(void* dev_ptr, long long n_doubles, void* stream, void* user);
int magi_hmc_set_global(magi_handle* h, long long n_chains_total, magi_allreduce_fn allreduce, void* user);
/* In-library collectives of a multi-rank run (NCCL over NVLink, resolved from libnccl.so.2 at run time).  Rank 0 calls
 * magi_nccl_unique_id and distributes the MAGI_NCCL_ID_BYTES bytes by any host-side means; every rank then calls magi_comm_init
 * (collective).  A host that already has an ncclComm_t on the handle's device hands it over with magi_comm_attach instead (the
 * host keeps ownership).  With a communicator set and no callback given to magi_hmc_set_global, magi_hmc_run sums the warm-up's
 * pooled window statistics with ncclAllReduce on its stream; magi_hmc_allgather_draws gathers the retained draws of all ranks:
 * out_dev is [world][n_stored][n_chains][n_cols] on the device, all ranks holding equally many chains and iterations. */
#define MAGI_NCCL_ID_BYTES 128
int magi_nccl_unique_id(char* id /* MAGI_NCCL_ID_BYTES */);
int magi_comm_init(magi_handle* h, const char* id, int rank, int world);
int magi_comm_attach(magi_handle* h, void* nccl_comm, int rank, int world);
int magi_comm_warmup(magi_handle* h, void* stream);
int magi_hmc_allgather_draws(magi_handle* h, double* out_dev, void* stream);
int magi_hmc_run(magi_handle* h, int n_iter, int n_leapfrog, int adapt, double target_accept, int store_draws, void* stream);
/* The same sampler with NUTS trajectories (multinomial sampling, generalised U-turn criterion; at most 2^max_depth - 1 leapfrog steps
 * per transition) instead of static ones: the batched counterpart of run_nuts_sampler's kernel, Trajectory{MultinomialTS}(Leapfrog,
 * GeneralisedNoUTurn) (src/samplers.jl:158-160).  All chains grow their trees in lock-step (doubling j = 2^j batched gradient evaluations);
 * chains whose tree has stopped wait.  magi_nuts_get_stats: mean tree depth / leapfrog steps per transition of every chain. */
int magi_nuts_run(magi_handle* h, int n_iter, int max_depth, int adapt, double target_accept, int store_draws, void* stream);
int magi_nuts_get_stats(magi_handle* h, double* mean_depth /* n_chains */, double* mean_leapfrog /* n_chains */);
int magi_hmc_reset_stats(magi_handle* h);
int magi_hmc_get_state(magi_handle* h, double* params /* P x n_chains or NULL */, double* ll /* n_chains or NULL */);
/* draws: [n_stored][n_chains][k + D + 1] = (theta, sigma, lp) like solve_magi's theta / sigma / lp (src/MagiJl.jl:633-771) */
int magi_hmc_get_draws(magi_handle* h, double* out, long long max_iters, long long* n_stored);
/* X draws: vec(X) (n*D doubles, time fastest) of the first n_chains_x chains at every thin-th kept iteration -- x_sampled of
 * solve_magi's result (src/MagiJl.jl:633-771).  magi_hmc_store_x after magi_hmc_init, before the kept iterations;
 * magi_hmc_get_x_draws fills out[n_stored][n_chains_x][n*D]. */
int magi_hmc_store_x(magi_handle* h, int n_chains_x, int thin);
int magi_hmc_get_x_draws(magi_handle* h, double* out, long long max_draws, long long* n_stored, int* n_chains_x);
int magi_hmc_draws_dev(magi_handle* h, void** ptr_dev, long long* n_stored, int* n_chains, int* n_cols);
int magi_hmc_get_stats(magi_handle* h, double* accept_rate, double* step_size, int* n_divergent, double* xmean /* nD x n_chains */,
                       double* minv /* P */);
long long magi_hmc_grad_evals(magi_handle* h);

/* introspection used by bench.py / tests: number of kernel launches issued by this handle so far */
long long magi_launch_count(const magi_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* MAGI_B200_H */
