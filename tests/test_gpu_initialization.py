"""GP hyper-parameter initialisation: device NLML against the oracle's restatement of
src/initialization.jl:72-176, and the Nelder-Mead search on top of it."""
import numpy as np
import pytest

from oracle import magi_oracle as mo
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel,n", [("matern52", 41), ("rbf", 25), ("matern52", 180)])
def test_nlml_matches_oracle(pkg, kernel, n):
    rng = np.random.default_rng(n)
    t = np.linspace(0.0, 20.0, n)
    y = H.fn_truth(t)[:, 0] + 0.2 * rng.normal(size=n)
    y[::7] = np.nan
    cands = np.log(np.array([[2.0, 1.5, 0.2], [0.5, 3.0, 0.5], [10.0, 0.3, 0.05], [1.0, 1.0, 1.0]]))
    got = pkg.initialization.negative_log_marginal_likelihood_batched(cands, y, t, kernel, jitter=1e-6)
    ref = np.array([mo.negative_log_marginal_likelihood(c, y, t, mo.KERNEL_IDS[kernel], 1e-6) for c in cands])
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-9
    assert pkg.initialization.negative_log_marginal_likelihood(np.array([np.inf, 0.0, 0.0]), y, t, kernel) == np.inf
    assert np.all(np.isinf(pkg.initialization.negative_log_marginal_likelihood_batched(cands, np.full(n, np.nan), t, kernel)))


def test_optimize_recovers_generating_hyperparameters(pkg):
    """Draw y from a GP with known (variance, lengthscale, sigma): the optimiser must land near them and must not be
    worse than the truth in NLML."""
    rng = np.random.default_rng(3)
    n = 120
    t = np.linspace(0.0, 30.0, n)
    true = np.array([2.0, 1.5, 0.3])
    K = mo.kernel_matrix(mo.MATERN52, t, true[0], true[1]) + 1e-9 * np.eye(n)
    y = np.linalg.cholesky(K) @ rng.normal(size=n) + true[2] * rng.normal(size=n)
    x0 = pkg.initialization.initial_guess(y, t)
    opt = pkg.initialization.optimize_gp_hyperparameters(y, t, "matern52", x0, jitter=1e-6, iterations=200)
    f_opt = pkg.initialization.negative_log_marginal_likelihood(np.log(opt), y, t, "matern52")
    f_true = pkg.initialization.negative_log_marginal_likelihood(np.log(true), y, t, "matern52")
    assert f_opt <= f_true + 1e-6
    assert np.all(np.abs(np.log(opt) - np.log(true)) < np.log(2.5)), opt


def test_solve_magi_without_phi_and_sigma(pkg):
    """test/runtests.jl:74-116 (unknown sigma, no phi): runs end to end with estimated hyper-parameters."""
    t = np.linspace(0.0, 20.0, 161)
    truth = H.fn_truth(t)
    rng = np.random.default_rng(5)
    y = np.full_like(truth, np.nan)
    y[::4] = truth[::4] + 0.2 * rng.normal(size=truth[::4].shape)
    res = pkg.solve_magi(y, t, pkg.fn_system(), dict(niterHmc=60, burninRatio=0.5, nChains=32, nLeapfrog=10, thetaInit=np.array([0.5, 0.5, 2.0])))
    assert res["phi"].shape == (2, 2) and np.all(res["phi"] > 0) and np.all(np.isfinite(res["phi"]))
    assert res["sigma"].shape == (30, 32, 2) and np.all(np.isfinite(res["lp"]))
    assert np.all(np.abs(res["sigma"][0, 0] - 0.2) < 0.3)
