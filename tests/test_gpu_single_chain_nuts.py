"""The single-chain drop-in (magi_logdensity_and_gradient) driven the way the reference drives it: run_nuts_sampler
(src/samplers.jl:114-194) calls it once per leapfrog step from one host task through wrappers that assert a finite value and
gradient (:53-63).  Here the host NUTS loop of the package does the same through the C ABI, on the reference's own end-to-end
test problem (test/runtests.jl:11-43: FN, t = 0:0.5:5); plus solve_magi's x_sampled."""
import warnings

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _problem(pkg):
    t = np.arange(0.0, 5.0 + 1e-9, 0.5)
    truth = H.fn_truth(t)
    rng = np.random.default_rng(123)
    y = truth + rng.normal(size=truth.shape) * np.array([0.25, 0.35])
    phi = np.array([[2.0, 1.0], [1.5, 2.0]])
    tg = pkg.MagiTarget.from_config(y, t, phi, pkg.fn_system(), np.array([0.3, 0.3]), bandsize=10, jitter=1e-6)
    p0 = np.concatenate([truth.reshape(-1, order="F"), [0.2, 0.2, 3.0], np.log([0.3, 0.3])])
    return tg, p0, t, y, phi


def test_host_nuts_drives_the_single_chain_entry(pkg):
    tg, p0, t, y, phi = _problem(pkg)
    l0 = tg.launch_count()
    chain, stats = pkg.run_nuts_sampler(tg, p0, n_samples=500, n_adapts=250, target_accept_ratio=0.8, initial_step_size=0.05, seed=1)
    assert chain is not None and chain.shape == (250, tg.dimension()) and np.all(np.isfinite(chain))
    n_grad = sum(s["n_leapfrog"] for s in stats)
    assert tg.launch_count() - l0 >= n_grad                                  # every leapfrog step was one call of the CUDA path
    acc = np.mean([s["accept_stat"] for s in stats])
    assert 0.55 < acc < 0.99, acc
    assert np.mean([s["divergent"] for s in stats]) < 0.05
    th = chain[:, 22:25].mean(axis=0); sg = np.exp(chain[:, 25:]).mean(axis=0)
    # the batched on-device sampler on the same posterior: means agree within Monte Carlo error of a 250-draw chain
    res = pkg.solve_magi(y, t, pkg.fn_system(), dict(niterHmc=800, burninRatio=0.5, bandSize=10, stepSizeFactor=0.005, phi=phi,
                                                      sigmaInit=np.array([0.3, 0.3]), nChains=256, nLeapfrog=25, seed=4))
    th_b, sd_b = res["theta"].mean(axis=(0, 1)), res["theta"].std(axis=(0, 1))
    assert np.all(np.abs(th - th_b) < 1.0 * sd_b), (th, th_b, sd_b)
    assert np.all(np.abs(sg - res["sigma"].mean(axis=(0, 1))) < 1.0 * res["sigma"].std(axis=(0, 1)))


def test_host_nuts_failure_modes_match_the_reference(pkg):
    tg, p0, *_ = _problem(pkg)
    with pytest.raises(AssertionError, match="dimension mismatch"):         # samplers.jl:125
        pkg.run_nuts_sampler(tg, p0[:-1], n_samples=4, n_adapts=2)
    bad = p0.copy(); bad[3] = np.nan                                          # (-Inf, zeros) from interface.jl:222-226 -> the wrapper's assert aborts the run
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        chain, stats = pkg.run_nuts_sampler(tg, bad, n_samples=4, n_adapts=2)
    assert chain is None and stats is None and any("NUTS" in str(x.message) for x in w)
    with pytest.raises(AssertionError, match="finite"):
        pkg.logdensity_and_gradient_func_wrapper(tg, bad)
    # a diverging trajectory (huge step) is flagged, not fatal: the energy error is finite or +Inf, the state stays put
    chain, stats = pkg.run_nuts_sampler(tg, p0, n_samples=6, n_adapts=0, initial_step_size=50.0, seed=2)
    assert chain is None or (np.all(np.isfinite(chain)) and any(s["divergent"] for s in stats))


def test_solve_magi_returns_x_sampled(pkg):
    """src/MagiJl.jl:633-771: theta S x k, x_sampled S x n x D, sigma S x D, lp S (per chain here)."""
    tg, p0, t, y, phi = _problem(pkg)
    cfg = dict(niterHmc=120, burninRatio=0.5, bandSize=10, stepSizeFactor=0.005, phi=phi, sigmaInit=np.array([0.3, 0.3]), nChains=32,
               nLeapfrog=10, seed=3, xChains=5)
    res = pkg.solve_magi(y, t, pkg.fn_system(), cfg)
    n, D = y.shape
    assert res["theta"].shape == (60, 32, 3) and res["x_sampled"].shape == (60, 5, n, D) and np.all(np.isfinite(res["x_sampled"]))
    assert np.allclose(res["x_sampled"].mean(axis=0), res["x_mean"][:5], rtol=1e-12, atol=1e-12)     # the same draws the running mean saw
    thin = pkg.solve_magi(y, t, pkg.fn_system(), dict(cfg, xThin=4, xChains=2))
    assert thin["x_sampled"].shape == (15, 2, n, D)
    assert np.array_equal(thin["x_sampled"], res["x_sampled"][::4, :2])        # same seed, same chains: every 4th kept draw
    one = pkg.solve_magi(y, t, pkg.fn_system(), dict(cfg, nChains=1))
    assert one["theta"].shape == (60, 3) and one["x_sampled"].shape == (60, n, D) and one["sigma"].shape == (60, 2) and one["lp"].shape == (60,)
