"""On-device batched HMC and solve_magi: statistical checks in the spirit of
test/runtests.jl:57-220 (posterior mean of theta within 0.5 and of sigma within 0.3 of the truth, FN, 11 time points)."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _fn_data(seed=123):
    t = np.arange(0.0, 5.0 + 1e-9, 0.5)                       # test/runtests.jl:17-20
    truth = H.fn_truth(t)
    rng = np.random.default_rng(seed)
    sig = np.array([0.25, 0.35])
    return t, truth + rng.normal(size=truth.shape) * sig, np.array([0.2, 0.2, 3.0]), sig


def test_posterior_means_match_cpu_oracle_sampler(pkg):
    """north_star: posterior means must agree within Monte Carlo error.  The comparator is an independent CPU HMC whose
    gradient is the oracle's C restatement of the reference (tests/cpu_hmc.py), on the reference's own end-to-end test
    problem (test/runtests.jl:11-43: FN, t = 0:0.5:5, sigma = [0.25, 0.35])."""
    from oracle import magi_oracle as mo
    from tests import cpu_hmc
    t, y, th_true, sig_true = _fn_data()
    phi = np.array([[2.0, 1.0], [1.5, 2.0]])
    cfg = dict(niterHmc=1200, burninRatio=0.5, bandSize=20, stepSizeFactor=0.005, phi=phi, sigmaInit=np.array([0.3, 0.3]),
               nChains=256, nLeapfrog=25, seed=1)
    res = pkg.solve_magi(y, t, pkg.fn_system(), cfg)
    assert res["theta"].shape == (600, 256, 3) and res["sigma"].shape == (600, 256, 2) and res["lp"].shape == (600, 256)
    assert np.all(np.isfinite(res["lp"]))
    assert 0.6 < np.median(res["stats"]["accept_rate"]) < 0.99
    s = pkg.diagnostics.summarize(res["theta"][:, :64], names=["a", "b", "c"])
    assert all(r["rhat"] < 1.1 for r in s), s
    covs = [mo.calculate_gp_covariances(mo.MATERN52, phi[:, d], t, 10, jitter=1e-6) for d in range(2)]
    tgt = mo.make_target(y, covs, mo.MODEL_FN, [0.3, 0.3], (1.0, 1.0, 1.0), False)
    p0 = np.concatenate([H.fn_truth(t).reshape(-1, order="F"), th_true, np.log(sig_true)])
    P0 = p0[None, :] + 0.01 * np.random.default_rng(0).normal(size=(16, len(p0)))
    minv = res["stats"]["inverse_metric"]
    draws, acc = cpu_hmc.cpu_hmc(tgt, P0, 1600, 600, float(np.median(res["stats"]["step_size"])), 25, minv=minv, seed=5)
    assert acc > 0.6
    th_cpu = draws[:, :, 22:25].mean(axis=(0, 1)); sg_cpu = np.exp(draws[:, :, 25:]).mean(axis=(0, 1))
    th_gpu = res["theta"].mean(axis=(0, 1)); sg_gpu = res["sigma"].mean(axis=(0, 1))
    th_sd = res["theta"].std(axis=(0, 1)); sg_sd = res["sigma"].std(axis=(0, 1))
    # Monte Carlo error of the 16-chain CPU run dominates: allow 0.25 posterior standard deviations
    assert np.all(np.abs(th_gpu - th_cpu) < 0.25 * th_sd), (th_gpu, th_cpu, th_sd)
    assert np.all(np.abs(sg_gpu - sg_cpu) < 0.25 * sg_sd), (sg_gpu, sg_cpu, sg_sd)


def test_solve_magi_recovers_truth_on_discretised_fn(pkg):
    """The statistical check of test/runtests.jl:108,115 on a problem discretised the way run_scripts/fn_example.jl does
    it (observations on a sub-grid of a finer time grid; MAGI's derivative constraint is biased on coarse grids: with the
    11-point grid of runtests.jl both this sampler and the CPU oracle sampler sit at theta ~ (0.83, 0.48, 1.27))."""
    t = np.linspace(0.0, 20.0, 321)
    truth = H.fn_truth(t)
    rng = np.random.default_rng(11)
    y = np.full_like(truth, np.nan)
    y[::8] = truth[::8] + 0.2 * rng.normal(size=truth[::8].shape)
    cfg = dict(niterHmc=1500, burninRatio=0.5, bandSize=20, stepSizeFactor=0.005, phi=np.array([[2.0, 1.0], [1.5, 2.0]]),
               sigmaInit=np.array([0.2, 0.2]), nChains=128, nLeapfrog=40, seed=2, thetaInit=np.array([0.5, 0.5, 2.0]))
    res = pkg.solve_magi(y, t, pkg.fn_system(), cfg)
    th = res["theta"].mean(axis=(0, 1)); sg = res["sigma"].mean(axis=(0, 1))
    assert abs(th[0] - 0.2) < 0.5 and abs(th[1] - 0.2) < 0.8 and abs(th[2] - 3.0) < 0.8, th
    assert np.all(np.abs(sg - 0.2) < 0.3), sg
    assert np.max(np.abs(res["x_mean"].mean(axis=0) - truth)) < 0.8


def test_solve_magi_fixed_sigma(pkg):
    """test/runtests.jl:121-182: with :sigma and :phi given, sigma is not sampled and every output row equals the input."""
    t, y, th_true, sig_true = _fn_data()
    cfg = dict(niterHmc=200, burninRatio=0.5, phi=np.array([[2.0, 1.0], [1.5, 2.0]]), sigma=sig_true, nChains=64, nLeapfrog=20)
    res = pkg.solve_magi(y, t, pkg.fn_system(), cfg)
    assert res["target"].dimension() == 11 * 2 + 3
    assert res["sigma"].shape == (100, 64, 2) and np.allclose(res["sigma"], sig_true[None, None, :])
    assert np.all(np.isfinite(res["theta"])) and np.all(np.isfinite(res["lp"]))


def test_hmc_is_invariant_to_sharding(pkg):
    """Philox streams are keyed by the global chain id: two half-size runs with offsets reproduce one full run bit for bit."""
    prob = H.make_problem(n=21, T=5.0, b=8, n_chains=16, seed=4, obs_every=2)
    def run(params, offset):
        tg = H.cuda_target(pkg, prob)
        chain, st = pkg.run_hmc_sampler(tg, params, n_samples=12, n_adapts=0, initial_step_size=0.002, n_leapfrog=5, seed=7, chain_id_offset=offset)
        return chain
    full = run(prob["params"], 0)
    a, b = run(prob["params"][:8], 0), run(prob["params"][8:], 8)
    assert np.array_equal(full[:, :8], a) and np.array_equal(full[:, 8:], b)


def test_hmc_is_reproducible_bit_for_bit(pkg):
    """Same seed, same draws -- including the warm-up (dual averaging, pooled windowed metric): the kinetic energy and the
    window statistics are reduced in a fixed order (no atomics), at a size where several warps and blocks contribute."""
    prob = H.make_problem(n=201, T=20.0, b=20, n_chains=8, seed=11, obs_every=5)
    rng = np.random.default_rng(3)
    params = np.repeat(prob["params"], 24, axis=0) + 1e-3 * rng.normal(size=(192, prob["params"].shape[1]))
    def run():
        tg = H.cuda_target(pkg, prob)
        chain, st = pkg.run_hmc_sampler(tg, params, n_samples=40, n_adapts=30, initial_step_size=0.002, n_leapfrog=5, seed=5)
        return chain, st
    c1, s1 = run()
    c2, s2 = run()
    assert np.array_equal(c1, c2)
    assert np.array_equal(s1["step_size"], s2["step_size"]) and np.array_equal(s1["inverse_metric"], s2["inverse_metric"])


def test_hmc_warmup_is_invariant_to_sharding(pkg):
    """Two "ranks" (two handles driven by two threads on one GPU) with the window all-reduce callback reproduce a one-rank
    run bit for bit, warm-up included: the pooled metric is reduced over fixed slices of GLOBAL chain ids in a fixed order
    (magi_hmc_set_global, include/magi_b200.h).  The callback here sums the two ranks' buffers through the host."""
    import threading
    import torch
    from manifold_constrained_gaussian_process_inference_b200 import _lib
    prob = H.make_problem(n=41, T=8.0, b=6, n_chains=8, seed=13, obs_every=2)
    rng = np.random.default_rng(5)
    params = np.repeat(prob["params"], 16, axis=0) + 1e-3 * rng.normal(size=(128, prob["params"].shape[1]))
    kw = dict(n_samples=60, n_adapts=45, initial_step_size=0.002, n_leapfrog=5, seed=9)
    full, st_full = pkg.run_hmc_sampler(H.cuda_target(pkg, prob), params, n_chains_total=128, **kw)

    barrier = threading.Barrier(2)
    slots = [None, None]

    def make_cb(rank):
        def cb(ptr, n, stream, user):
            class _Holder:
                pass
            hld = _Holder()
            hld.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
            t = torch.as_tensor(hld, device="cuda:0")
            torch.cuda.synchronize()
            slots[rank] = t.cpu()
            barrier.wait()
            total = slots[0] + slots[1]
            barrier.wait()
            t.copy_(total.to("cuda:0"))
            torch.cuda.synchronize()
            return 0
        return _lib.ALLREDUCE_FN(cb)

    out = [None, None]

    def worker(rank):
        tg = H.cuda_target(pkg, prob)
        chain, st = pkg.run_hmc_sampler(tg, params[64 * rank:64 * (rank + 1)], chain_id_offset=64 * rank, n_chains_total=128,
                                        window_allreduce=make_cb(rank), **kw)
        out[rank] = (chain, st)

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(2)]
    for t in threads: t.start()
    for t in threads: t.join(timeout=120)
    assert out[0] is not None and out[1] is not None
    assert np.array_equal(full[:, :64], out[0][0]) and np.array_equal(full[:, 64:], out[1][0])
    assert np.array_equal(st_full["inverse_metric"], out[0][1]["inverse_metric"])
    assert np.array_equal(st_full["step_size"][:64], out[0][1]["step_size"]) and np.array_equal(st_full["step_size"][64:], out[1][1]["step_size"])


def test_hmc_energy_conservation_and_reversibility_proxy(pkg):
    """With a tiny step the acceptance probability must be ~1 (the leapfrog integrates the gradient the kernel returns)."""
    prob = H.make_problem(n=41, T=8.0, b=6, n_chains=32, seed=2)
    tg = H.cuda_target(pkg, prob)
    chain, st = pkg.run_hmc_sampler(tg, prob["params"], n_samples=10, n_adapts=0, initial_step_size=1e-4, n_leapfrog=10, seed=3)
    assert np.all(st["accept_rate"] > 0.98), st["accept_rate"]
    assert st["grad_evals"] == 32 * (1 + 10 * 10)


def test_batched_nuts_matches_static_hmc_and_is_reproducible(pkg):
    """The batched NUTS transition (magi_nuts_run; run_nuts_sampler's kernel, src/samplers.jl:158-160) on the reference's own test
    problem (test/runtests.jl:11-43): same posterior as the static-trajectory sampler within Monte Carlo error, sane tree
    statistics, bit-identical reruns, and results that do not depend on how the chains are batched."""
    t, y, th_true, sig_true = _fn_data()
    phi = np.array([[2.0, 1.0], [1.5, 2.0]])
    tg = pkg.MagiTarget.from_config(y, t, phi, pkg.fn_system(), np.array([0.3, 0.3]), bandsize=10, jitter=1e-6)
    p0 = np.concatenate([H.fn_truth(t).reshape(-1, order="F"), th_true, np.log(sig_true)])
    P0 = p0[None, :] + 0.01 * np.random.default_rng(0).normal(size=(128, len(p0)))
    kw = dict(n_samples=500, n_adapts=250, target_accept_ratio=0.8, initial_step_size=0.01, seed=7, max_tree_depth=8)
    chain, st = pkg.run_hmc_sampler(tg, P0, **kw)
    assert chain.shape == (250, 128, 6) and np.all(np.isfinite(chain))
    assert 0.6 < np.median(st["accept_rate"]) < 0.97, np.median(st["accept_rate"])
    assert 1.0 <= st["tree_depth"].mean() <= 8.0 and np.all(st["n_leapfrog_mean"] >= 1.0)
    assert np.mean(st["n_divergent"]) < 5
    s = pkg.diagnostics.summarize(chain[:, :64, :3], names=["a", "b", "c"])
    assert all(r["rhat"] < 1.1 for r in s), s
    chain2, _ = pkg.run_hmc_sampler(tg, P0, **kw)
    assert np.array_equal(chain, chain2)                                   # bit-identical rerun
    sub, _ = pkg.run_hmc_sampler(tg, P0[:32], n_chains_total=128, **kw)    # a shard of the batch: the trees of other chains do not matter ...
    # ... but the pooled metric of the warm-up does: with the global chain count set, the shard needs the other shard's statistics,
    # which a single rank does not have -- so compare the kept draws of a run WITHOUT adaptation
    kw0 = dict(kw, n_adapts=0, n_samples=60)
    full0, _ = pkg.run_hmc_sampler(tg, P0, **kw0)
    sub0, _ = pkg.run_hmc_sampler(tg, P0[32:64], chain_id_offset=32, **kw0)
    assert np.array_equal(full0[:, 32:64], sub0)
    ref, str_ = pkg.run_hmc_sampler(tg, P0, n_samples=800, n_adapts=400, target_accept_ratio=0.8, initial_step_size=0.01, seed=8, n_leapfrog=25)
    for j in range(5):
        a, b = chain[:, :, j], ref[:, :, j]
        assert abs(a.mean() - b.mean()) < 0.2 * b.std(), (j, a.mean(), b.mean(), b.std())
        assert 0.75 < a.std() / b.std() < 1.3, (j, a.std(), b.std())
