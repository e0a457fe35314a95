"""Shared builders for the tests: seeded synthetic MAGI problems (SURVEY.md section 8(d) shapes) evaluated by the
oracle (test infrastructure) and by the CUDA path through the C ABI."""
from __future__ import annotations

import numpy as np

from oracle import magi_oracle as mo

LL_RTOL = 1e-10           # north_star: 1e-10 relative on the log density ...
GRAD_RTOL = 1e-10         # ... and on the gradient, measured against max(|g_i|, 1e-3 * ||g||_inf) (SURVEY.md section 7:
GRAD_FLOOR = 1e-3         # individual elements carry ~1e-12 relative rounding from row-level cancellation in C~ x)


def fn_truth(tvec, theta=(0.2, 0.2, 3.0), x0=(-1.0, 1.0)):
    from scipy.integrate import solve_ivp
    if len(tvec) == 1:
        return np.array([x0], dtype=np.float64)
    a, b, c = theta
    f = lambda t, u: [c * (u[0] - u[0] ** 3 / 3 + u[1]), -(u[0] - a + b * u[1]) / c]
    sol = solve_ivp(f, (tvec[0], tvec[-1]), x0, t_eval=tvec, rtol=1e-8, atol=1e-10)
    return sol.y.T


def lv_truth(tvec, theta=(1.5, 1.0, 3.0, 1.0), x0=(1.0, 1.0)):
    from scipy.integrate import solve_ivp
    al, be, de, ga = theta
    f = lambda t, u: [al * u[0] - be * u[0] * u[1], de * u[0] * u[1] - ga * u[1]]
    sol = solve_ivp(f, (tvec[0], tvec[-1]), x0, t_eval=tvec, rtol=1e-8, atol=1e-10)
    return sol.y.T


def make_problem(model="fn", n=41, T=8.0, b=6, n_chains=5, seed=0, obs_every=3, phis=None, kernel=mo.MATERN52,
                 jitter=1e-6, beta=(1.0, 1.0, 1.0), sigma_fixed=False, noise=0.2, setup_mode="reference_order"):
    """Returns dict(target_oracle, params[n_chains, P], meta...) for a synthetic problem."""
    rng = np.random.default_rng(seed)
    tvec = np.linspace(0.0, T, n)
    if model == "fn":
        mid, truth, th_true = mo.MODEL_FN, fn_truth(tvec), np.array([0.2, 0.2, 3.0])
        phis = phis or [(2.0, 1.5), (1.0, 2.0)]
    elif model == "lv":
        mid, truth, th_true = mo.MODEL_LV, lv_truth(tvec), np.array([1.5, 1.0, 3.0, 1.0])
        phis = phis or [(1.0, 1.5), (1.0, 1.5)]
    elif model == "hes1":
        mid, th_true = mo.MODEL_HES1, np.array([0.022, 0.3, 0.031, 0.028, 0.5, 20.0, 0.3])
        truth = np.stack([1.5 + np.sin(tvec / 3), 2.0 + np.cos(tvec / 3), 3.0 + 0.5 * np.sin(tvec / 2)], axis=1)
        phis = phis or [(2.0, 1.5), (1.0, 2.0), (1.5, 1.0)]
    else:
        raise ValueError(model)
    D = truth.shape[1]
    b = min(b, n - 1)
    covs = [mo.calculate_gp_covariances(kernel, phis[d], tvec, b, complexity=2, jitter=jitter, setup_mode=setup_mode) for d in range(D)]
    Y = np.full((n, D), np.nan)
    Y[::obs_every] = truth[::obs_every] + noise * rng.normal(size=truth[::obs_every].shape)
    sigma_init = np.full(D, noise)
    tgt = mo.make_target(Y, covs, mid, sigma_init, beta, sigma_fixed)
    P = mo.dimension(tgt)
    params = np.zeros((n_chains, P))
    for c in range(n_chains):
        X = truth + 0.1 * rng.normal(size=truth.shape)
        th = th_true * np.exp(0.1 * rng.normal(size=th_true.shape))
        parts = [X.reshape(-1, order="F"), th]
        if not sigma_fixed:
            parts.append(np.log(noise) + 0.1 * rng.normal(size=D))
        params[c] = np.concatenate(parts)
    return dict(target=tgt, params=params, tvec=tvec, covs=covs, Y=Y, model=model, model_id=mid, sigma_init=sigma_init,
                beta=beta, sigma_fixed=sigma_fixed, b=b, phis=phis, kernel=kernel, jitter=jitter)


def cuda_target(pkg, prob, **kw):
    """The CUDA MagiTarget for an oracle problem, fed the oracle's band tables (parity boundary (i))."""
    t = prob["target"]
    covs = []
    for g in prob["covs"]:
        c = pkg.GPCov(phi=np.asarray(g.phi, dtype=np.float64), tvec=np.asarray(g.tvec, dtype=np.float64), bandsize=int(g.bandsize),
                      CinvBand=np.asarray(g.CinvBand, dtype=np.float64), mphiBand=np.asarray(g.mphiBand, dtype=np.float64),
                      KinvBand=np.asarray(g.KinvBand, dtype=np.float64))
        covs.append(c)
    name = {mo.MODEL_FN: "fn", mo.MODEL_HES1: "hes1", mo.MODEL_LV: "lv", mo.MODEL_HES1LOG: "hes1log", mo.MODEL_HES1LOG_FIXG: "hes1log_fixg",
            mo.MODEL_HES1LOG_FIXF: "hes1log_fixf", mo.MODEL_HIV: "hiv", mo.MODEL_PTRANS: "ptrans", mo.MODEL_L96: "lorenz96"}[t.model.model_id]
    return pkg.MagiTarget(np.asarray(t.yobs, dtype=np.float64), covs, pkg.get_ode_system(name, t.n_dims), np.asarray(t.sigma_init, dtype=np.float64),
                          list(t.prior_temperature), t.n_times, t.n_dims, t.n_params_ode, t.sigma_is_fixed, **kw)


def oracle_batched(prob, params=None):
    params = prob["params"] if params is None else params
    lls, grads = [], []
    for p in params:
        ll, g = mo.logdensity_and_gradient(prob["target"], p)
        lls.append(ll)
        grads.append(g)
    return np.array(lls), np.array(grads)


def assert_parity(ll, grad, ll_ref, grad_ref, what=""):
    ll, ll_ref = np.asarray(ll), np.asarray(ll_ref)
    fin = np.isfinite(ll_ref)
    assert np.array_equal(np.isfinite(ll), fin), what + ": finite pattern of ll differs"
    assert np.all(ll[~fin] == ll_ref[~fin]), what + ": non-finite ll values differ"
    rel = np.abs(ll[fin] - ll_ref[fin]) / np.maximum(np.abs(ll_ref[fin]), 1e-300)
    assert rel.size == 0 or rel.max() <= LL_RTOL, "%s: ll relative error %.3e > %.1e" % (what, rel.max(), LL_RTOL)
    if grad is not None:
        grad, grad_ref = np.atleast_2d(grad), np.atleast_2d(grad_ref)
        assert np.array_equal(np.isnan(grad), np.isnan(grad_ref)), what + ": NaN pattern of grad differs"
        ok = ~np.isnan(grad_ref)
        scale = np.maximum(np.abs(grad_ref), GRAD_FLOOR * np.nanmax(np.abs(grad_ref), axis=1, keepdims=True))
        err = np.where(ok, np.abs(grad - grad_ref) / np.where(scale > 0, scale, 1.0), 0.0)
        assert err.max() <= GRAD_RTOL, "%s: gradient error %.3e > %.1e" % (what, err.max(), GRAD_RTOL)
    return float(rel.max()) if rel.size else 0.0


def fn_n3_golden():
    """(ll, grad[11], rtol): the FN N=3 case of test/test_likelihoods.jl:18-59 in long double (tests/golden/reference_known_answers.json);
    the single constant both the oracle pins and the GPU parity tests compare with."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_known_answers.json")) as f:
        g = json.load(f)["restated_fn_n3_values"]
    return float(g["ll"]), np.array(g["grad"]), float(g["rtol"])
