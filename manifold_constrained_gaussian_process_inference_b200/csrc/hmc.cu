// On-device batched HMC: the caller either side of the hot path.
// The reference drives ONE chain with AdvancedHMC NUTS from the host (run_nuts_sampler, src/samplers.jl:114-194:
// DiagEuclideanMetric, Leapfrog, StanHMCAdaptor(MassMatrixAdaptor, StepSizeAdaptor(delta))), calling
// logdensity_and_gradient once per leapfrog step.  Doing that for thousands of chains through PCIe costs ~0.5 ms per step
// against tens of microseconds of kernel, so the state stays resident in HBM and a transition is a short sequence of
// launches on one stream:
//   momentum refresh (Philox, keyed by global chain id => results do not depend on how chains are sharded over GPUs)
//   L x [ p += eps/2 g ; q += eps Minv p ; gradient (K1/K2) ; p += eps/2 g ]   (the two half kicks are fused)
//   Metropolis accept/reject per chain, Nesterov dual averaging of eps per chain (Stan / AdvancedHMC defaults:
//   gamma 0.05, t0 10, kappa 0.75, mu = log(10 eps0)), pooled diagonal metric from cross-chain variances in Stan-style
//   doubling windows.  Static trajectory length (batched NUTS would diverge per chain; a later row).
#include <cmath>
#include <cstdio>
#include "magi_internal.cuh"

namespace magi {

#define HCK(call, what) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_error(e__, what); } while (0)

struct HmcState {
    int n_chains = 0, P = 0, n_draw_cols = 0;
    long long chain_offset = 0;
    unsigned long long seed = 0;
    long long iter = 0;               // transitions done so far (RNG counter)
    long long grad_evals = 0;
    double *q = nullptr, *p = nullptr, *g = nullptr, *ll = nullptr;        // current position / momentum / gradient / log density
    double *q0 = nullptr, *g0 = nullptr, *ll0 = nullptr, *h0 = nullptr;    // start of the trajectory
    double *minv = nullptr;                                              // diagonal inverse metric (shared by all chains)
    double *eps = nullptr, *da = nullptr;                                // per-chain step size and dual-averaging state [5]
    double *acc_sum = nullptr;                                           // per-chain sum of acceptance probabilities
    int *n_div = nullptr;
    double *wsum = nullptr, *wsq = nullptr;                              // pooled window accumulators [P]
    double *wpart = nullptr;                                             // [kWinSlices][2][P] per-slice sums of the current window
    long long n_chains_total = 0;                                        // chains of ALL ranks (slices are defined on global chain ids)
    int (*allreduce)(void*, long long, void*, void*) = nullptr;          // optional in-place sum over ranks (magi_hmc_set_global)
    void* allreduce_user = nullptr;
    long long wcount = 0;
    double *draws = nullptr; long long draws_cap = 0, n_draws = 0;       // [n_draws][n_chains][k + D + 1]
    double *xsum = nullptr; long long xsum_count = 0;                    // per-chain running sum of vec(X) over kept draws
    // X draws (solve_magi's x_sampled, src/MagiJl.jl:633-771): vec(X) of the first x_chains chains at every x_thin-th kept draw
    double *xdraws = nullptr; long long xdraws_cap = 0, n_xdraws = 0;
    int x_chains = 0, x_thin = 1;
    long long acc_count = 0;
    // ---- batched NUTS (magi_nuts_run): per-chain tree state; allocated at the first NUTS run ----
    double *ql = nullptr, *pl = nullptr, *gl = nullptr, *qr = nullptr, *pr = nullptr, *gr = nullptr;   // end points of the tree
    double *qprop = nullptr, *gprop = nullptr, *llprop = nullptr;      // multinomial sample of the whole tree
    double *qsub = nullptr, *gsub = nullptr, *llsub = nullptr;         // ... of the subtree being built
    double *rho = nullptr, *srho = nullptr;                            // sums of the momenta: tree, subtree
    double *pck = nullptr, *rck = nullptr;                             // [max_depth][n_chains][P] checkpoints (momentum, cumulative sum)
    double *lsw = nullptr, *slsw = nullptr, *sum_acc = nullptr;        // log sum of weights: tree, subtree; sum of min(1, exp(-dH))
    int *tflags = nullptr, *tdir = nullptr, *nleap = nullptr, *tdepth = nullptr, *n_active = nullptr;
    long long *depth_sum = nullptr, *leap_sum = nullptr;               // statistics over the iterations since magi_hmc_reset_stats
    int nuts_max_depth = 0;
};
constexpr int kTreeDone = 1, kTreeSubInvalid = 2, kTreeDiverged = 4;

// ---- Philox4x32-10 counter RNG ----
__device__ __forceinline__ void philox4x32(unsigned int c[4], unsigned int k0, unsigned int k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const unsigned int n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ double u01(unsigned int a, unsigned int b) {   // (0, 1)
    const unsigned long long v = ((unsigned long long)a << 21) ^ (unsigned long long)b;   // 53 bits
    return ((double)(v & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
}
// Philox key of a chain: (seed, global chain id) hashed TOGETHER (splitmix64 of the seed, plus the chain id, hashed again), so
// that the streams of (seed s, chain c) and (seed s ^ k, chain c ^ k) are unrelated -- a key of `seed ^ chain` made runs with
// nearby seeds reuse one permuted set of streams.  The counter words carry (pair | draw kind, stream, iteration).
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ unsigned long long chain_key(unsigned long long seed, long long chain) {
    return splitmix64(splitmix64(seed) + (unsigned long long)chain);
}
// two standard normals for (chain, iteration, pair index)
__device__ __forceinline__ void normal_pair(unsigned long long seed, long long chain, long long iter, unsigned int pair, unsigned int stream,
                                            double& z0, double& z1) {
    unsigned int c[4] = {pair, stream, (unsigned int)iter, (unsigned int)(iter >> 32)};
    const unsigned long long key = chain_key(seed, chain);
    philox4x32(c, (unsigned int)key, (unsigned int)(key >> 32));
    const double u1 = u01(c[0], c[1]), u2 = u01(c[2], c[3]);
    const double r = sqrt(-2.0 * log(u1));
    double s, co;
    sincospi(2.0 * u2, &s, &co);
    z0 = r * co; z1 = r * s;
}

// momentum refresh + snapshot of the trajectory start + first half kick + drift
__global__ void hmc_begin_kernel(HmcState s, int P) {
    const long long c = blockIdx.x;
    const double eps = s.eps[c];
    double kin = 0.0;
    const size_t base = (size_t)c * P;
    for (int i2 = blockIdx.y * blockDim.x + threadIdx.x; 2 * i2 < P; i2 += gridDim.y * blockDim.x) {
        double z[2];
        normal_pair(s.seed, c + s.chain_offset, s.iter, (unsigned int)i2, 0u, z[0], z[1]);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = 2 * i2 + u;
            if (i < P) {
                const double mi = s.minv[i];
                double p = z[u] * rsqrt(mi);                 // p ~ N(0, M), M = 1 / Minv
                kin += 0.5 * p * p * mi;
                const double q = s.q[base + i], g = s.g[base + i];
                s.q0[base + i] = q; s.g0[base + i] = g;
                p += 0.5 * eps * g;
                s.p[base + i] = p;
                s.q[base + i] = q + eps * mi * p;
            }
        }
    }
    // ordered reduction (no atomics: the Hamiltonian, hence every accept / reject decision, is reproducible bit for bit)
    __shared__ double sh[8];
    for (int o = 16; o > 0; o >>= 1) kin += __shfl_xor_sync(0xffffffffu, kin, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = kin;
    __syncthreads();
    if (threadIdx.x == 0) {
        double k = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) k += sh[i];
        s.h0[c] += k;
    }
}

// h0[c] = -ll[c] (before the kinetic energy is accumulated); ll0 snapshot
__global__ void hmc_prep_kernel(HmcState s) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < s.n_chains) { s.ll0[c] = s.ll[c]; s.h0[c] = -s.ll[c]; }
}

// between two gradient evaluations: p += eps g (two fused half kicks); q += eps Minv p
__global__ void hmc_kick_drift_kernel(HmcState s, int P) {
    const long long c = blockIdx.x;
    const double eps = s.eps[c];
    const size_t base = (size_t)c * P;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < P; i += gridDim.y * blockDim.x) {
        const double p = s.p[base + i] + eps * s.g[base + i];
        s.p[base + i] = p;
        s.q[base + i] += eps * s.minv[i] * p;
    }
}

struct HmcFinish {
    int adapt, store, n_draw_cols, nD, K, D, sigma_is_fixed;
    double delta, mu_scale;
    int accumulate_window, accumulate_x, store_x;
    int nuts;            // 1: the transition was a NUTS tree (state, acceptance statistic and divergence come from the tree)
};

// last half kick, Hamiltonian, accept/reject, dual averaging, draw storage (one block per chain)
__global__ void __launch_bounds__(256) hmc_finish_kernel(HmcState s, int P, HmcFinish f) {
    __shared__ double sh[8];
    __shared__ int s_accept;
    const long long c = blockIdx.x;
    const size_t base = (size_t)c * P;
    const double eps = s.eps[c];
    double kin = 0.0;
    if (f.nuts) {                                   // the tree's multinomial sample becomes the state
        for (int i = threadIdx.x; i < P; i += blockDim.x) { s.q[base + i] = s.qprop[base + i]; s.g[base + i] = s.gprop[base + i]; }
    } else {
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            const double p = s.p[base + i] + 0.5 * eps * s.g[base + i];
            kin += 0.5 * p * p * s.minv[i];
        }
    }
    for (int o = 16; o > 0; o >>= 1) kin += __shfl_xor_sync(0xffffffffu, kin, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = kin;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a;
        bool count_div;
        if (f.nuts) {
            s.ll[c] = s.llprop[c];
            a = s.nleap[c] > 0 ? s.sum_acc[c] / s.nleap[c] : 0.0;          // acceptance statistic of the tree (Stan / AdvancedHMC)
            count_div = (s.tflags[c] & kTreeDiverged) != 0;
            s_accept = 1;
            s.depth_sum[c] += s.tdepth[c]; s.leap_sum[c] += s.nleap[c];
        } else {
            double k1 = 0.0;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) k1 += sh[i];
            const double h1 = -s.ll[c] + k1;
            a = exp(s.h0[c] - h1);
            const bool divergent = !isfinite(h1) || !isfinite(a) && !(s.h0[c] - h1 > 0);
            if (!isfinite(h1)) a = 0.0;
            if (a > 1.0 || (isinf(a) && s.h0[c] - h1 > 0)) a = 1.0;
            if (isnan(a)) a = 0.0;
            unsigned int cc[4] = {1u, 2u, (unsigned int)s.iter, (unsigned int)(s.iter >> 32)};     // stream 2: the accept uniform
            const unsigned long long key = chain_key(s.seed, c + s.chain_offset);
            philox4x32(cc, (unsigned int)key, (unsigned int)(key >> 32));
            const double u = u01(cc[0], cc[1]);
            s_accept = (u < a) ? 1 : 0;
            count_div = divergent || (s.h0[c] - h1) < -1000.0;
        }
        s.acc_sum[c] += a;
        if (count_div) s.n_div[c] += 1;
        if (f.adapt) {
            // Nesterov dual averaging (Hoffman & Gelman 2014; AdvancedHMC/Stan defaults)
            double* da = s.da + (size_t)c * 5;     // m, Hbar, log_eps_bar, mu, (unused)
            const double gamma = 0.05, t0 = 10.0, kappa = 0.75;
            const double m = da[0] + 1.0;
            const double eta = 1.0 / (m + t0);
            const double hbar = (1.0 - eta) * da[1] + eta * (f.delta - a);
            const double log_eps = da[3] - sqrt(m) / gamma * hbar;
            const double w = pow(m, -kappa);
            const double log_eps_bar = w * log_eps + (1.0 - w) * da[2];
            da[0] = m; da[1] = hbar; da[2] = log_eps_bar;
            s.eps[c] = exp(log_eps);
        }
    }
    __syncthreads();
    const bool acc = s_accept != 0;
    if (!acc) {
        for (int i = threadIdx.x; i < P; i += blockDim.x) { s.q[base + i] = s.q0[base + i]; s.g[base + i] = s.g0[base + i]; }
        if (threadIdx.x == 0) s.ll[c] = s.ll0[c];
    }
    __syncthreads();
    if (f.accumulate_x) for (int i = threadIdx.x; i < f.nD; i += blockDim.x) s.xsum[(size_t)c * f.nD + i] += s.q[base + i];
    if (f.store_x && c < s.x_chains) {
        double* xo = s.xdraws + ((size_t)s.n_xdraws * s.x_chains + c) * f.nD;
        for (int i = threadIdx.x; i < f.nD; i += blockDim.x) xo[i] = s.q[base + i];
    }
    if (f.store && threadIdx.x < f.n_draw_cols) {
        double* out = s.draws + ((size_t)s.n_draws * s.n_chains + c) * f.n_draw_cols;
        const int j = threadIdx.x;
        double v;
        if (j < f.K) v = s.q[base + f.nD + j];                                        // theta
        else if (j < f.K + f.D) {                                                     // sigma = exp(clamped log sigma) (MagiJl.jl:696)
            if (f.sigma_is_fixed) v = nan("");
            else v = exp(fmin(fmax(s.q[base + f.nD + j], -15.0), 15.0));
        } else v = s.ll[c];                                                           // lp
        out[j] = v;
    }
}

// Pooled window statistics without atomics, identical for every sharding of the chains: the GLOBAL chain ids are cut into
// kWinSlices fixed slices; a slice's running sums add this iteration's chains in ascending order; at the end of a window the
// slices are (optionally summed over ranks -- every slice lives on one rank, the others add exact zeros -- and) merged in
// slice order.
constexpr int kWinSlices = 64;
__global__ void hmc_window_partial_kernel(HmcState s, int P, double* __restrict__ part) {      // grid (ceil(P/128), kWinSlices)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const long long per = (s.n_chains_total + kWinSlices - 1) / kWinSlices;
    long long c0 = (long long)blockIdx.y * per - s.chain_offset, c1 = c0 + per;
    if (c0 < 0) c0 = 0;
    if (c1 > s.n_chains) c1 = s.n_chains;
    double a = 0.0, b = 0.0;
    for (long long c = c0; c < c1; ++c) { const double q = s.q[(size_t)c * P + i]; a += q; b += q * q; }
    if (c1 > c0) {
        part[(size_t)blockIdx.y * 2 * P + i] += a;
        part[(size_t)blockIdx.y * 2 * P + P + i] += b;
    }
}
__global__ void hmc_window_merge_kernel(HmcState s, int P, double* __restrict__ part) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    double a = 0.0, b = 0.0;
    for (int k = 0; k < kWinSlices; ++k) {
        a += part[(size_t)k * 2 * P + i]; b += part[(size_t)k * 2 * P + P + i];
        part[(size_t)k * 2 * P + i] = 0.0; part[(size_t)k * 2 * P + P + i] = 0.0;
    }
    s.wsum[i] = a; s.wsq[i] = b;
}

__global__ void hmc_window_finish_kernel(HmcState s, int P, double count, int reset_only) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) {
        if (!reset_only && count > 1.0) {
            const double mean = s.wsum[i] / count;
            double var = (s.wsq[i] - count * mean * mean) / (count - 1.0);
            if (!(var > 0.0) || !isfinite(var)) var = 1.0;
            // Stan's shrinkage towards unit metric; count here is chains x iterations, so it is a light touch
            const double nn = count;
            s.minv[i] = (nn / (nn + 5.0)) * var + 1e-3 * (5.0 / (nn + 5.0));
        }
        s.wsum[i] = 0.0; s.wsq[i] = 0.0;
    }
}

__global__ void hmc_restart_da_kernel(HmcState s, int finalize) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < s.n_chains) {
        double* da = s.da + (size_t)c * 5;
        if (finalize) { if (da[0] > 0) s.eps[c] = exp(da[2]); }       // use the averaged step size after warm-up
        else { da[0] = 0.0; da[1] = 0.0; da[2] = 0.0; da[3] = log(10.0 * s.eps[c]); }
    }
}

__global__ void fill_double_kernel(double* p, size_t nel, double v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nel; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}


// ------------------------------------------------------------------------------------------------------------------------
// Batched NUTS (run_nuts_sampler's trajectory, src/samplers.jl:158-160: Trajectory{MultinomialTS}(Leapfrog, GeneralisedNoUTurn)).
// AdvancedHMC is not under the reference tree: the transition is restated from its published algorithm (Hoffman & Gelman
// 2014; Betancourt 2017) in the ITERATIVE form (no recursion: a subtree of 2^j leaves is built leaf by leaf, the U-turn
// checks inside it use O(depth) checkpoints of (momentum, cumulative momentum sum) addressed by the bits of the leaf index;
// Phan, Pradhan, Jankowiak 2019).  All chains grow their trees in lock-step -- depth j costs 2^j batched gradient
// evaluations for every chain still growing; a chain whose tree has stopped is masked and waits -- so one chain never
// changes another chain's result and the draws do not depend on how the chains are sharded.
// Per-chain random numbers: Philox counter words (kind, index, iteration): 3 = direction of doubling `index`, 4 = multinomial
// choice at leaf `index` of the transition, 5 = choice between the old tree and the new subtree at doubling `index`.
// ------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double chain_uniform(const HmcState& s, long long c, unsigned int kind, unsigned int index) {
    unsigned int cc[4] = {kind, index, (unsigned int)s.iter, (unsigned int)(s.iter >> 32)};
    const unsigned long long key = chain_key(s.seed, c + s.chain_offset);
    philox4x32(cc, (unsigned int)key, (unsigned int)(key >> 32));
    return u01(cc[0], cc[1]);
}
__device__ __forceinline__ double log_add_exp(double a, double b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    const double m = fmax(a, b);
    return m + log(exp(a - m) + exp(b - m));
}
__device__ __forceinline__ double block_sum_256(double v, double* sh) {      // ordered: bit-reproducible
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sh[i];
    return r;
}

// momentum refresh, H0, a tree of one node (the current state), direction of the first doubling
__global__ void __launch_bounds__(256) nuts_begin_kernel(HmcState s, int P) {
    __shared__ double sh[8];
    const long long c = blockIdx.x;
    const size_t base = (size_t)c * P;
    double kin = 0.0;
    for (int i2 = threadIdx.x; 2 * i2 < P; i2 += blockDim.x) {
        double z[2];
        normal_pair(s.seed, c + s.chain_offset, s.iter, (unsigned int)i2, 0u, z[0], z[1]);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = 2 * i2 + u;
            if (i < P) {
                const double mi = s.minv[i], p = z[u] * rsqrt(mi), q = s.q[base + i], g = s.g[base + i];
                kin += 0.5 * p * p * mi;
                s.p[base + i] = p;
                s.ql[base + i] = q; s.qr[base + i] = q; s.qprop[base + i] = q;
                s.pl[base + i] = p; s.pr[base + i] = p; s.rho[base + i] = p;
                s.gl[base + i] = g; s.gr[base + i] = g; s.gprop[base + i] = g;
                s.srho[base + i] = 0.0;
            }
        }
    }
    kin = block_sum_256(kin, sh);
    if (threadIdx.x == 0) {
        s.h0[c] = -s.ll[c] + kin;
        s.llprop[c] = s.ll[c];
        s.lsw[c] = 0.0; s.slsw[c] = -INFINITY; s.sum_acc[c] = 0.0;
        s.nleap[c] = 0; s.tdepth[c] = 0; s.tflags[c] = 0;
        s.tdir[c] = chain_uniform(s, c, 3u, 0u) < 0.5 ? 1 : -1;
    }
}

// first half of a leapfrog step of the moving state of every chain whose subtree is still growing
__global__ void nuts_pre_kernel(HmcState s, int P) {
    const long long c = blockIdx.x;
    if (s.tflags[c] & (kTreeDone | kTreeSubInvalid)) return;
    const double eps = s.tdir[c] * s.eps[c];
    const size_t base = (size_t)c * P;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < P; i += gridDim.y * blockDim.x) {
        const double p = s.p[base + i] + 0.5 * eps * s.g[base + i];
        s.p[base + i] = p;
        s.q[base + i] += eps * s.minv[i] * p;
    }
}

// second half kick of leaf `leaf` of the current doubling; energy error, multinomial choice inside the subtree, checkpoints
// and U-turn checks inside the subtree
__global__ void __launch_bounds__(256) nuts_post_kernel(HmcState s, int P, int leaf) {
    __shared__ double sh[8];
    __shared__ int s_take;
    const long long c = blockIdx.x;
    if (s.tflags[c] & (kTreeDone | kTreeSubInvalid)) return;
    const size_t base = (size_t)c * P, NP = (size_t)s.n_chains * P;
    const double eps = s.tdir[c] * s.eps[c];
    double kin = 0.0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const double p = s.p[base + i] + 0.5 * eps * s.g[base + i];
        s.p[base + i] = p;
        s.srho[base + i] += p;
        kin += 0.5 * p * p * s.minv[i];
    }
    kin = block_sum_256(kin, sh);
    if (threadIdx.x == 0) {
        double dh = (-s.ll[c] + kin) - s.h0[c];
        if (!isfinite(dh)) dh = INFINITY;
        const double w = -dh;
        s.sum_acc[c] += dh <= 0.0 ? 1.0 : exp(-dh);
        const int nl = s.nleap[c] + 1;
        s.nleap[c] = nl;
        const double nw = log_add_exp(s.slsw[c], w);
        s_take = (w > -INFINITY && log(chain_uniform(s, c, 4u, (unsigned int)nl)) < w - nw) ? 1 : 0;     // uniform over the subtree's leaves by weight
        s.slsw[c] = nw;
        if (dh > 1000.0) s.tflags[c] |= kTreeDiverged | kTreeSubInvalid;
    }
    __syncthreads();
    if (s_take) {
        for (int i = threadIdx.x; i < P; i += blockDim.x) { s.qsub[base + i] = s.q[base + i]; s.gsub[base + i] = s.g[base + i]; }
        if (threadIdx.x == 0) s.llsub[c] = s.ll[c];
    }
    const int idx_max = __popc(leaf >> 1);
    if ((leaf & 1) == 0) {                              // even leaf: becomes the checkpoint of its level
        double* pk = s.pck + (size_t)idx_max * NP + base;
        double* rk = s.rck + (size_t)idx_max * NP + base;
        for (int i = threadIdx.x; i < P; i += blockDim.x) { pk[i] = s.p[base + i]; rk[i] = s.srho[base + i]; }
    } else {                                            // odd leaf: closes `ntrail` nested subtrees; U-turn check of each
        const int ntrail = __ffs(~leaf) - 1;            // number of trailing one bits
        bool turning = false;
        for (int k = idx_max; k >= idx_max - ntrail + 1; --k) {
            const double* pk = s.pck + (size_t)k * NP + base;
            const double* rk = s.rck + (size_t)k * NP + base;
            double a = 0.0, b = 0.0;
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const double r = s.srho[base + i] - rk[i] + pk[i];          // momentum sum of the nested subtree
                a += s.minv[i] * pk[i] * r;
                b += s.minv[i] * s.p[base + i] * r;
            }
            a = block_sum_256(a, sh); b = block_sum_256(b, sh);
            turning |= (a <= 0.0) || (b <= 0.0);
        }
        if (threadIdx.x == 0 && turning) s.tflags[c] |= kTreeSubInvalid;
    }
}

// end of doubling `depth`: merge the subtree into the tree (biased progressive sampling), U-turn check of the whole tree,
// next direction; chains that go on are counted in *n_active
__global__ void __launch_bounds__(256) nuts_merge_kernel(HmcState s, int P, int depth, int max_depth) {
    __shared__ double sh[8];
    __shared__ int s_acc, s_dir;
    const long long c = blockIdx.x;
    int fl = s.tflags[c];
    if (fl & kTreeDone) return;
    if (fl & kTreeSubInvalid) { if (threadIdx.x == 0) s.tflags[c] = fl | kTreeDone; return; }     // the subtree is discarded
    const size_t base = (size_t)c * P;
    const int v = s.tdir[c];
    if (threadIdx.x == 0) s_acc = log(chain_uniform(s, c, 5u, (unsigned int)depth)) < s.slsw[c] - s.lsw[c] ? 1 : 0;
    __syncthreads();
    double* qe = v > 0 ? s.qr : s.ql; double* pe = v > 0 ? s.pr : s.pl; double* ge = v > 0 ? s.gr : s.gl;
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        if (s_acc) { s.qprop[base + i] = s.qsub[base + i]; s.gprop[base + i] = s.gsub[base + i]; }
        qe[base + i] = s.q[base + i]; pe[base + i] = s.p[base + i]; ge[base + i] = s.g[base + i];      // the moving state is the new end point
        const double r = s.rho[base + i] + s.srho[base + i];
        s.rho[base + i] = r;
        s.srho[base + i] = 0.0;
        a += s.minv[i] * s.pl[base + i] * r;
        b += s.minv[i] * s.pr[base + i] * r;
    }
    a = block_sum_256(a, sh); b = block_sum_256(b, sh);
    if (threadIdx.x == 0) {
        if (s_acc) s.llprop[c] = s.llsub[c];
        s.lsw[c] = log_add_exp(s.lsw[c], s.slsw[c]);
        s.slsw[c] = -INFINITY;
        s.tdepth[c] = depth + 1;
        if (a <= 0.0 || b <= 0.0 || depth + 1 >= max_depth) { s.tflags[c] = fl | kTreeDone; s_dir = 0; }
        else {
            s_dir = chain_uniform(s, c, 3u, (unsigned int)(depth + 1)) < 0.5 ? 1 : -1;
            s.tdir[c] = s_dir;
            atomicAdd(s.n_active, 1);
        }
    }
    __syncthreads();
    if (s_dir != 0) {                                   // the next subtree starts from the end point in the new direction
        const double* qs = s_dir > 0 ? s.qr : s.ql; const double* ps = s_dir > 0 ? s.pr : s.pl; const double* gs = s_dir > 0 ? s.gr : s.gl;
        for (int i = threadIdx.x; i < P; i += blockDim.x) { s.q[base + i] = qs[base + i]; s.p[base + i] = ps[base + i]; s.g[base + i] = gs[base + i]; }
    }
}

void hmc_free(magi_handle* h) {
    HmcState* s = (HmcState*)h->hmc;
    if (!s) return;
    double* ptrs[] = {s->q, s->p, s->g, s->ll, s->q0, s->g0, s->ll0, s->h0, s->minv, s->eps, s->da, s->acc_sum, s->wsum, s->wsq, s->wpart, s->draws, s->xsum, s->xdraws};
    for (double* p : ptrs) if (p) cudaFree(p);
    double* tree[] = {s->ql, s->pl, s->gl, s->qr, s->pr, s->gr, s->qprop, s->gprop, s->llprop, s->qsub, s->gsub, s->llsub, s->rho, s->srho, s->pck, s->rck,
                      s->lsw, s->slsw, s->sum_acc};
    for (double* p : tree) if (p) cudaFree(p);
    int* ti[] = {s->tflags, s->tdir, s->nleap, s->tdepth, s->n_active};
    for (int* p : ti) if (p) cudaFree(p);
    if (s->depth_sum) cudaFree(s->depth_sum);
    if (s->leap_sum) cudaFree(s->leap_sum);
    if (s->n_div) cudaFree(s->n_div);
    delete s;
    h->hmc = nullptr;
}

}  // namespace magi

using namespace magi;

extern "C" int magi_hmc_init(magi_handle* h, int n_chains, const double* params0, unsigned long long seed, double step_size0,
                             long long chain_id_offset) {
    if (!h || !params0 || n_chains <= 0 || !(step_size0 > 0)) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_hmc_init: bad argument");
    HCK(cudaSetDevice(h->device), "cudaSetDevice");
    hmc_free(h);
    HmcState* s = new HmcState();
    h->hmc = s;
    const int P = h->P;
    s->n_chains = n_chains; s->P = P; s->seed = seed; s->chain_offset = chain_id_offset;
    s->n_draw_cols = h->K + h->D + 1;
    const size_t NP = (size_t)n_chains * P;
    double** big[] = {&s->q, &s->p, &s->g, &s->q0, &s->g0};
    for (double** b : big) HCK(cudaMalloc(b, sizeof(double) * NP), "cudaMalloc hmc state");
    double** small[] = {&s->ll, &s->ll0, &s->h0, &s->eps, &s->acc_sum};
    for (double** b : small) HCK(cudaMalloc(b, sizeof(double) * n_chains), "cudaMalloc hmc per-chain");
    HCK(cudaMalloc(&s->da, sizeof(double) * 5 * n_chains), "cudaMalloc da");
    HCK(cudaMalloc(&s->n_div, sizeof(int) * n_chains), "cudaMalloc n_div");
    HCK(cudaMalloc(&s->minv, sizeof(double) * P), "cudaMalloc minv");
    HCK(cudaMalloc(&s->wsum, sizeof(double) * P), "cudaMalloc wsum");
    HCK(cudaMalloc(&s->wsq, sizeof(double) * P), "cudaMalloc wsq");
    HCK(cudaMalloc(&s->wpart, sizeof(double) * 2 * P * kWinSlices), "cudaMalloc window slice sums");
    HCK(cudaMemsetAsync(s->wpart, 0, sizeof(double) * 2 * P * kWinSlices, h->stream), "memset");
    s->n_chains_total = n_chains;
    HCK(cudaMalloc(&s->xsum, sizeof(double) * (size_t)n_chains * h->n * h->D), "cudaMalloc xsum");
    cudaStream_t st = h->stream;
    HCK(cudaMemcpyAsync(s->q, params0, sizeof(double) * NP, cudaMemcpyHostToDevice, st), "H2D initial state");
    fill_double_kernel<<<64, 256, 0, st>>>(s->minv, P, 1.0);
    fill_double_kernel<<<64, 256, 0, st>>>(s->eps, n_chains, step_size0);
    HCK(cudaMemsetAsync(s->acc_sum, 0, sizeof(double) * n_chains, st), "memset");
    HCK(cudaMemsetAsync(s->n_div, 0, sizeof(int) * n_chains, st), "memset");
    HCK(cudaMemsetAsync(s->wsum, 0, sizeof(double) * P, st), "memset");
    HCK(cudaMemsetAsync(s->wsq, 0, sizeof(double) * P, st), "memset");
    HCK(cudaMemsetAsync(s->xsum, 0, sizeof(double) * (size_t)n_chains * h->n * h->D, st), "memset");
    hmc_restart_da_kernel<<<(n_chains + 255) / 256, 256, 0, st>>>(*s, 0);
    h->launches += 3;
    int rc = eval_dev(h, n_chains, s->q, P, s->ll, s->g, st);
    if (rc) return rc;
    s->grad_evals += n_chains;
    HCK(cudaStreamSynchronize(st), "hmc init sync");
    return MAGI_OK;
}

// Multi-rank runs: the total number of chains over all ranks (the pooled metric is a statistic of ALL chains) and an in-place
// sum-over-ranks callback for the window statistics, called on the sampler's stream at the end of every adaptation window.
// With both set, warm-up and sampling are bit-identical to a one-rank run over the same global chains whenever the shard
// boundaries fall on slice boundaries (n_chains_total / 64 chains per slice).  Call after magi_hmc_init.
extern "C" int magi_hmc_set_global(magi_handle* h, long long n_chains_total, int (*allreduce)(void*, long long, void*, void*), void* user) {
    if (!h || !h->hmc) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_hmc_set_global: call magi_hmc_init first");
    HmcState* s = (HmcState*)h->hmc;
    if (n_chains_total < s->n_chains + s->chain_offset) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_hmc_set_global: n_chains_total is smaller than this rank's last chain id");
    s->n_chains_total = n_chains_total; s->allreduce = allreduce; s->allreduce_user = user;
    return MAGI_OK;
}

// Runs n_iter transitions of n_leapfrog steps.  adapt != 0: warm-up (dual averaging towards target_accept and windowed
// metric adaptation over these n_iter iterations).  store_draws != 0: appends (theta, sigma, lp) of every chain per iteration.
// tree state of the batched NUTS, allocated at the first run (or when a deeper tree is asked for)
static int nuts_alloc(magi_handle* h, HmcState* s, int max_depth) {
    const size_t NP = (size_t)s->n_chains * s->P, nc = (size_t)s->n_chains;
    if (!s->ql) {
        double** big[] = {&s->ql, &s->pl, &s->gl, &s->qr, &s->pr, &s->gr, &s->qprop, &s->gprop, &s->qsub, &s->gsub, &s->rho, &s->srho};
        for (double** b : big) HCK(cudaMalloc(b, sizeof(double) * NP), "cudaMalloc NUTS tree state");
        double** small[] = {&s->llprop, &s->llsub, &s->lsw, &s->slsw, &s->sum_acc};
        for (double** b : small) HCK(cudaMalloc(b, sizeof(double) * nc), "cudaMalloc NUTS per-chain");
        int** si[] = {&s->tflags, &s->tdir, &s->nleap, &s->tdepth};
        for (int** b : si) HCK(cudaMalloc(b, sizeof(int) * nc), "cudaMalloc NUTS per-chain");
        HCK(cudaMalloc(&s->n_active, sizeof(int)), "cudaMalloc");
        HCK(cudaMalloc(&s->depth_sum, sizeof(long long) * nc), "cudaMalloc");
        HCK(cudaMalloc(&s->leap_sum, sizeof(long long) * nc), "cudaMalloc");
        HCK(cudaMemset(s->depth_sum, 0, sizeof(long long) * nc), "memset");
        HCK(cudaMemset(s->leap_sum, 0, sizeof(long long) * nc), "memset");
    }
    if (max_depth > s->nuts_max_depth) {
        if (s->pck) cudaFree(s->pck);
        if (s->rck) cudaFree(s->rck);
        s->pck = s->rck = nullptr;
        HCK(cudaMalloc(&s->pck, sizeof(double) * NP * max_depth), "cudaMalloc NUTS checkpoints");
        HCK(cudaMalloc(&s->rck, sizeof(double) * NP * max_depth), "cudaMalloc NUTS checkpoints");
        s->nuts_max_depth = max_depth;
    }
    (void)h;
    return MAGI_OK;
}

// n_leapfrog >= 1, max_depth = 0: static trajectories; max_depth >= 1: NUTS trees of at most 2^max_depth - 1 leapfrog steps
static int hmc_run_impl(magi_handle* h, int n_iter, int n_leapfrog, int max_depth, int adapt, double target_accept, int store_draws, void* stream_) {
    HCK(cudaSetDevice(h->device), "cudaSetDevice");
    HmcState* s = (HmcState*)h->hmc;
    cudaStream_t st = stream_ ? (cudaStream_t)stream_ : h->stream;
    const int P = s->P, nc = s->n_chains;
    if (max_depth > 0) { int rc = nuts_alloc(h, s, max_depth); if (rc) return rc; }
    if (store_draws) {
        const long long need = s->n_draws + n_iter;
        if (need > s->draws_cap) {
            double* nd = nullptr;
            HCK(cudaMalloc(&nd, sizeof(double) * (size_t)need * nc * s->n_draw_cols), "cudaMalloc draws");
            if (s->draws) { HCK(cudaMemcpyAsync(nd, s->draws, sizeof(double) * (size_t)s->n_draws * nc * s->n_draw_cols, cudaMemcpyDeviceToDevice, st), "copy draws"); HCK(cudaStreamSynchronize(st), "sync"); cudaFree(s->draws); }
            s->draws = nd; s->draws_cap = need;
        }
    }
    if (store_draws && s->x_chains > 0) {
        const long long need = s->n_xdraws + (n_iter + s->x_thin - 1) / s->x_thin + 1;
        if (need > s->xdraws_cap) {
            double* nd = nullptr;
            const size_t per = (size_t)s->x_chains * h->n * h->D;
            HCK(cudaMalloc(&nd, sizeof(double) * (size_t)need * per), "cudaMalloc X draws");
            if (s->xdraws) { HCK(cudaMemcpyAsync(nd, s->xdraws, sizeof(double) * (size_t)s->n_xdraws * per, cudaMemcpyDeviceToDevice, st), "copy X draws"); HCK(cudaStreamSynchronize(st), "sync"); cudaFree(s->xdraws); }
            s->xdraws = nd; s->xdraws_cap = need;
        }
    }
    // Stan-style windows over the warm-up: initial fast buffer, doubling slow windows, terminal fast buffer
    int init_buf = 75, term_buf = 50, base_win = 25;
    if (adapt && n_iter < 150) { init_buf = (int)(0.15 * n_iter); term_buf = (int)(0.10 * n_iter); base_win = n_iter - init_buf - term_buf; }
    int win_end = init_buf + base_win, win_size = base_win;
    if (adapt) { int rem = n_iter - term_buf - win_end; if (rem < 2 * win_size) win_end = n_iter - term_buf; }
    const dim3 egrid(nc, (P + 255) / 256 > 8 ? 8 : (P + 255) / 256);
    HmcFinish f;
    f.adapt = adapt; f.n_draw_cols = s->n_draw_cols; f.nD = h->n * h->D; f.K = h->K; f.D = h->D; f.sigma_is_fixed = h->sigma_is_fixed;
    f.delta = target_accept; f.mu_scale = 10.0;
    h->dispatch_chains = s->n_chains_total > 0 ? s->n_chains_total : 0;      // the same K1 variant on every rank of a sharded run
    struct DispatchReset { magi_handle* h; ~DispatchReset() { h->dispatch_chains = 0; } } dispatch_reset{h};
    f.nuts = max_depth > 0 ? 1 : 0;
    for (int it = 0; it < n_iter; ++it) {
        if (max_depth > 0) {
            nuts_begin_kernel<<<nc, 256, 0, st>>>(*s, P);
            h->launches++;
            for (int depth = 0; depth < max_depth; ++depth) {
                HCK(cudaMemsetAsync(s->n_active, 0, sizeof(int), st), "memset");
                for (int leaf = 0; leaf < (1 << depth); ++leaf) {
                    nuts_pre_kernel<<<egrid, 256, 0, st>>>(*s, P);
                    int rc = eval_dev(h, nc, s->q, P, s->ll, s->g, st);
                    if (rc) return rc;
                    s->grad_evals += nc;                     // (masked chains included: the batch is evaluated as a whole)
                    nuts_post_kernel<<<nc, 256, 0, st>>>(*s, P, leaf);
                    h->launches += 2;
                }
                nuts_merge_kernel<<<nc, 256, 0, st>>>(*s, P, depth, max_depth);
                h->launches++;
                int active = 0;
                HCK(cudaMemcpyAsync(&active, s->n_active, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H active chains");
                HCK(cudaStreamSynchronize(st), "nuts sync");
                if (active == 0) break;
            }
        } else {
        hmc_prep_kernel<<<(nc + 255) / 256, 256, 0, st>>>(*s);
        hmc_begin_kernel<<<dim3(nc, 1), 256, 0, st>>>(*s, P);      // one block per chain (ordered reduction of the kinetic energy)
        h->launches += 2;
        for (int l = 0; l < n_leapfrog; ++l) {
            int rc = eval_dev(h, nc, s->q, P, s->ll, s->g, st);
            if (rc) return rc;
            s->grad_evals += nc;
            if (l + 1 < n_leapfrog) { hmc_kick_drift_kernel<<<egrid, 256, 0, st>>>(*s, P); h->launches++; }
        }
        }
        const bool in_slow = adapt && it >= init_buf && it < n_iter - term_buf;
        f.store = store_draws; f.accumulate_window = in_slow ? 1 : 0; f.accumulate_x = store_draws ? 1 : 0;
        f.store_x = (store_draws && s->x_chains > 0 && (s->n_draws % s->x_thin) == 0) ? 1 : 0;
        hmc_finish_kernel<<<nc, 256, 0, st>>>(*s, P, f);
        h->launches++;
        if (in_slow) {
            hmc_window_partial_kernel<<<dim3((P + 127) / 128, kWinSlices), 128, 0, st>>>(*s, P, s->wpart);
            h->launches++;
        }
        s->iter++;
        if (f.store_x) s->n_xdraws++;
        if (store_draws) { s->n_draws++; s->xsum_count++; }
        s->acc_count++;
        if (in_slow) {
            s->wcount += s->n_chains_total;
            if (it + 1 == win_end) {
                // sum of the per-slice sums over the ranks, in place, on this stream: ncclAllReduce on the handle's communicator
                // (magi_comm_init / magi_comm_attach), or a host callback (magi_hmc_set_global) for transports other than NCCL
                if (s->allreduce) {
                    int rc = s->allreduce(s->wpart, (long long)kWinSlices * 2 * P, (void*)st, s->allreduce_user);
                    if (rc) return set_error(MAGI_ERR_CUDA, "magi_hmc_run: the window all-reduce callback failed");
                } else {
                    int rc = comm_allreduce_sum(h, s->wpart, (size_t)kWinSlices * 2 * P, st);
                    if (rc) return rc;
                }
                hmc_window_merge_kernel<<<(P + 127) / 128, 128, 0, st>>>(*s, P, s->wpart);
                h->launches++;
                hmc_window_finish_kernel<<<(P + 255) / 256, 256, 0, st>>>(*s, P, (double)s->wcount, 0);
                hmc_restart_da_kernel<<<(nc + 255) / 256, 256, 0, st>>>(*s, 0);
                h->launches += 2;
                s->wcount = 0;
                win_size *= 2;
                int next_end = win_end + win_size;
                if (n_iter - term_buf - next_end < 2 * win_size) next_end = n_iter - term_buf;
                win_end = next_end;
            }
        }
    }
    if (adapt && n_iter > 0) { hmc_restart_da_kernel<<<(nc + 255) / 256, 256, 0, st>>>(*s, 1); h->launches++; }
    HCK(cudaGetLastError(), "hmc kernels");
    HCK(cudaStreamSynchronize(st), "hmc sync");
    return MAGI_OK;
}

extern "C" int magi_hmc_run(magi_handle* h, int n_iter, int n_leapfrog, int adapt, double target_accept, int store_draws, void* stream_) {
    if (!h || !h->hmc || n_iter < 0 || n_leapfrog < 1) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_hmc_run: bad argument (call magi_hmc_init first)");
    return hmc_run_impl(h, n_iter, n_leapfrog, 0, adapt, target_accept, store_draws, stream_);
}

// The same sampler with NUTS trajectories (multinomial sampling, generalised U-turn criterion) of at most 2^max_depth - 1
// leapfrog steps instead of static ones: the batched counterpart of run_nuts_sampler's kernel (src/samplers.jl:158-160).
extern "C" int magi_nuts_run(magi_handle* h, int n_iter, int max_depth, int adapt, double target_accept, int store_draws, void* stream_) {
    if (!h || !h->hmc || n_iter < 0 || max_depth < 1 || max_depth > 12) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_nuts_run: bad argument (call magi_hmc_init first; 1 <= max_depth <= 12)");
    return hmc_run_impl(h, n_iter, 1, max_depth, adapt, target_accept, store_draws, stream_);
}

// mean tree depth and mean number of leapfrog steps per transition of every chain since magi_hmc_reset_stats (NUTS runs)
extern "C" int magi_nuts_get_stats(magi_handle* h, double* mean_depth, double* mean_leapfrog) {
    if (!h || !h->hmc) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_nuts_get_stats: no sampler state");
    HmcState* s = (HmcState*)h->hmc;
    if (!s->depth_sum) return set_error(MAGI_ERR_NOT_READY, "magi_nuts_get_stats: no NUTS run yet");
    HCK(cudaSetDevice(h->device), "cudaSetDevice");
    std::vector<long long> a(s->n_chains), b(s->n_chains);
    HCK(cudaMemcpy(a.data(), s->depth_sum, sizeof(long long) * s->n_chains, cudaMemcpyDeviceToHost), "D2H");
    HCK(cudaMemcpy(b.data(), s->leap_sum, sizeof(long long) * s->n_chains, cudaMemcpyDeviceToHost), "D2H");
    const double cnt = s->acc_count > 0 ? (double)s->acc_count : 1.0;
    for (int c = 0; c < s->n_chains; ++c) { if (mean_depth) mean_depth[c] = a[c] / cnt; if (mean_leapfrog) mean_leapfrog[c] = b[c] / cnt; }
    return MAGI_OK;
}

extern "C" int magi_hmc_get_state(magi_handle* h, double* params, double* ll) {
    if (!h || !h->hmc) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_hmc_get_state: no sampler state");
    HmcState* s = (HmcState*)h->hmc;
    HCK(cudaSetDevice(h->device), "cudaSetDevice");
    if (params) HCK(cudaMemcpy(params, s->q, sizeof(double) * (size_t)s->n_chains * s->P, cudaMemcpyDeviceToHost), "D2H state");
    if (ll) HCK(cudaMemcpy(ll, s->ll, sizeof(double) * s->n_chains, cudaMemcpyDeviceToHost), "D2H ll");
    return MAGI_OK;
}

// draws: [n_stored][n_chains][k + D + 1] = (theta, sigma, lp); returns the number of stored iterations through *n_stored
extern "C" int magi_hmc_get_draws(magi_handle* h, double* out, long long max_iters, long long* n_stored) {
    if (!h || !h->hmc) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_hmc_get_draws: no sampler state");
    HmcState* s = (HmcState*)h->hmc;
    HCK(cudaSetDevice(h->device), "cudaSetDevice");
    if (n_stored) *n_stored = s->n_draws;
    if (out && s->n_draws > 0) {
        const long long n = s->n_draws < max_iters ? s->n_draws : max_iters;
        HCK(cudaMemcpy(out, s->draws, sizeof(double) * (size_t)n * s->n_chains * s->n_draw_cols, cudaMemcpyDeviceToHost), "D2H draws");
    }
    return MAGI_OK;
}

// device pointer to the draw store (for an NCCL all-gather without a host round trip) and its geometry
extern "C" int magi_hmc_draws_dev(magi_handle* h, void** ptr, long long* n_stored, int* n_chains, int* n_cols) {
    if (!h || !h->hmc) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_hmc_draws_dev: no sampler state");
    HmcState* s = (HmcState*)h->hmc;
    if (ptr) *ptr = s->draws;
    if (n_stored) *n_stored = s->n_draws;
    if (n_chains) *n_chains = s->n_chains;
    if (n_cols) *n_cols = s->n_draw_cols;
    return MAGI_OK;
}

// End-of-run all-gather of the retained draws over the handle's communicator: out_dev receives [world][n_stored][n_chains][n_cols]
// (every rank must hold the same number of chains and of stored iterations); asynchronous on `stream`.
extern "C" int magi_hmc_allgather_draws(magi_handle* h, double* out_dev, void* stream) {
    if (!h || !h->hmc || !out_dev) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_hmc_allgather_draws: bad argument");
    HmcState* s = (HmcState*)h->hmc;
    HCK(cudaSetDevice(h->device), "cudaSetDevice");
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    const size_t cnt = (size_t)s->n_draws * s->n_chains * s->n_draw_cols;
    if (!h->nccl_comm || h->nccl_world <= 1) {
        HCK(cudaMemcpyAsync(out_dev, s->draws, sizeof(double) * cnt, cudaMemcpyDeviceToDevice, st), "copy draws");
        return MAGI_OK;
    }
    return comm_allgather(h, s->draws, out_dev, cnt, st);
}

// X draws: keep vec(X) (n*D doubles, time fastest) of the first n_chains_x chains at every thin-th kept iteration -- the x_sampled
// field of solve_magi's result (src/MagiJl.jl:633-771).  Call after magi_hmc_init and before the kept iterations.
extern "C" int magi_hmc_store_x(magi_handle* h, int n_chains_x, int thin) {
    if (!h || !h->hmc || n_chains_x < 0 || thin < 1) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_hmc_store_x: bad argument (call magi_hmc_init first)");
    HmcState* s = (HmcState*)h->hmc;
    if (s->n_xdraws > 0 && n_chains_x != s->x_chains) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_hmc_store_x: X draws of another chain count are already stored");
    s->x_chains = n_chains_x < s->n_chains ? n_chains_x : s->n_chains;
    s->x_thin = thin;
    return MAGI_OK;
}

// out: [n_stored][n_chains_x][n*D]; returns the number of stored X draws and the chain count through the pointers
extern "C" int magi_hmc_get_x_draws(magi_handle* h, double* out, long long max_draws, long long* n_stored, int* n_chains_x) {
    if (!h || !h->hmc) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_hmc_get_x_draws: no sampler state");
    HmcState* s = (HmcState*)h->hmc;
    HCK(cudaSetDevice(h->device), "cudaSetDevice");
    const long long ns = s->n_xdraws < max_draws ? s->n_xdraws : max_draws;
    if (n_stored) *n_stored = s->n_xdraws;
    if (n_chains_x) *n_chains_x = s->x_chains;
    if (out && ns > 0) HCK(cudaMemcpy(out, s->xdraws, sizeof(double) * (size_t)ns * s->x_chains * h->n * h->D, cudaMemcpyDeviceToHost), "D2H X draws");
    return MAGI_OK;
}

// per-chain statistics: mean acceptance probability, current step size, divergences; xmean = posterior mean of vec(X) per chain
extern "C" int magi_hmc_get_stats(magi_handle* h, double* accept_rate, double* step_size, int* n_divergent, double* xmean, double* minv) {
    if (!h || !h->hmc) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_hmc_get_stats: no sampler state");
    HmcState* s = (HmcState*)h->hmc;
    HCK(cudaSetDevice(h->device), "cudaSetDevice");
    const int nc = s->n_chains;
    if (accept_rate) {
        HCK(cudaMemcpy(accept_rate, s->acc_sum, sizeof(double) * nc, cudaMemcpyDeviceToHost), "D2H accept");
        for (int i = 0; i < nc; ++i) accept_rate[i] /= (double)(s->acc_count > 0 ? s->acc_count : 1);
    }
    if (step_size) HCK(cudaMemcpy(step_size, s->eps, sizeof(double) * nc, cudaMemcpyDeviceToHost), "D2H eps");
    if (n_divergent) HCK(cudaMemcpy(n_divergent, s->n_div, sizeof(int) * nc, cudaMemcpyDeviceToHost), "D2H ndiv");
    if (xmean) {
        const size_t nx = (size_t)nc * h->n * h->D;
        HCK(cudaMemcpy(xmean, s->xsum, sizeof(double) * nx, cudaMemcpyDeviceToHost), "D2H xsum");
        const double inv = 1.0 / (double)(s->xsum_count > 0 ? s->xsum_count : 1);
        for (size_t i = 0; i < nx; ++i) xmean[i] *= inv;
    }
    if (minv) HCK(cudaMemcpy(minv, s->minv, sizeof(double) * s->P, cudaMemcpyDeviceToHost), "D2H minv");
    return MAGI_OK;
}

extern "C" long long magi_hmc_grad_evals(magi_handle* h) {
    if (!h || !h->hmc) return -1;
    return ((HmcState*)h->hmc)->grad_evals;
}

// resets acceptance statistics and draw store (called between warm-up and sampling)
extern "C" int magi_hmc_reset_stats(magi_handle* h) {
    if (!h || !h->hmc) return set_error(MAGI_ERR_INVALID_ARGUMENT, "magi_hmc_reset_stats: no sampler state");
    HmcState* s = (HmcState*)h->hmc;
    HCK(cudaSetDevice(h->device), "cudaSetDevice");
    HCK(cudaMemset(s->acc_sum, 0, sizeof(double) * s->n_chains), "memset");
    HCK(cudaMemset(s->n_div, 0, sizeof(int) * s->n_chains), "memset");
    HCK(cudaMemset(s->xsum, 0, sizeof(double) * (size_t)s->n_chains * h->n * h->D), "memset");
    s->acc_count = 0; s->n_draws = 0; s->xsum_count = 0; s->n_xdraws = 0;
    if (s->depth_sum) { HCK(cudaMemset(s->depth_sum, 0, sizeof(long long) * s->n_chains), "memset"); HCK(cudaMemset(s->leap_sum, 0, sizeof(long long) * s->n_chains), "memset"); }
    return MAGI_OK;
}
