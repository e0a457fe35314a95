"""GP hyper-parameter initialisation (reference: src/initialization.jl).  The objective (negative log marginal
likelihood: covariance build, Cholesky, log-determinant, solve) is evaluated on the GPU for batches of candidates
(csrc/nlml_kernel.cu); the Nelder-Mead simplex search, which the reference delegates to Optim.jl 1.12 (not vendored), is
host logic restated here from its published algorithm (adaptive parameters of Gao & Han 2012 and the affine initial simplex,
Optim's defaults); its iterates are therefore parity-unpinned, only the objective is checked against the oracle."""
from __future__ import annotations

import numpy as np

from . import _lib


def negative_log_marginal_likelihood_batched(log_params, y_obs_dim, t_obs, kernel_type: str, jitter: float = 1e-6, device: int = 0):
    """NLML for a batch of candidates ``log_params`` (m, 3) = [log var, log len, log sigma] (src/initialization.jl:72-176).
    NaN observations are dropped (:91-100); no valid observation -> Inf (:93-96)."""
    L = _lib.lib()
    lp = np.ascontiguousarray(np.atleast_2d(log_params), dtype=np.float64)
    y = np.asarray(y_obs_dim, dtype=np.float64)
    t = np.asarray(t_obs, dtype=np.float64)
    ok = ~np.isnan(y)
    if not ok.any():
        return np.full(lp.shape[0], np.inf)
    ys, ts = np.ascontiguousarray(y[ok]), np.ascontiguousarray(t[ok])
    out = np.empty(lp.shape[0])
    kid = _lib.KERNEL_RBF if kernel_type == "rbf" else _lib.KERNEL_MATERN52          # unsupported types default to matern52 (:112-114)
    _lib.check(L.magi_gp_nlml_batched(kid, int(ys.shape[0]), _lib.as_dp(ts), _lib.as_dp(ys), float(jitter), int(lp.shape[0]),
                                      _lib.as_dp(lp), _lib.as_dp(out), int(device)))
    return out


def negative_log_marginal_likelihood(log_params, y_obs_dim, t_obs, kernel_type: str, jitter: float = 1e-6, device: int = 0) -> float:
    return float(negative_log_marginal_likelihood_batched(np.asarray(log_params)[None, :], y_obs_dim, t_obs, kernel_type, jitter, device)[0])


def _nelder_mead(f_batch, x0, iterations=100, g_tol=1e-8):
    """Nelder-Mead with Optim.jl's defaults: AffineSimplexer(a=0.025, b=0.5), AdaptiveParameters
    (alpha=1, beta=1+2/n, gamma=0.75-1/(2n), delta=1-1/n), convergence when the standard deviation of the simplex
    values drops below g_tol.  ``f_batch`` evaluates a (m, n) array of points."""
    x0 = np.asarray(x0, dtype=np.float64)
    n = x0.shape[0]
    alpha, beta, gamma, delta = 1.0, 1.0 + 2.0 / n, 0.75 - 1.0 / (2 * n), 1.0 - 1.0 / n
    simplex = np.tile(x0, (n + 1, 1))
    for i in range(n):
        simplex[i + 1, i] = (1.0 + 0.5) * x0[i] + 0.025
    fvals = f_batch(simplex)
    converged = False
    for it in range(iterations):
        order = np.argsort(fvals, kind="stable")
        simplex, fvals = simplex[order], fvals[order]
        if np.all(np.isfinite(fvals)) and np.sqrt(np.var(fvals) * n / (n + 1)) <= g_tol:
            converged = True
            break
        centroid = simplex[:n].mean(axis=0)
        xr = centroid + alpha * (centroid - simplex[n])
        xe = centroid + beta * (xr - centroid)
        xoc = centroid + gamma * (xr - centroid)
        xic = centroid - gamma * (xr - centroid)
        fr, fe, foc, fic = f_batch(np.stack([xr, xe, xoc, xic]))        # one launch evaluates every candidate of the step
        if fr < fvals[0]:
            if fe < fr: simplex[n], fvals[n] = xe, fe
            else: simplex[n], fvals[n] = xr, fr
        elif fr < fvals[n - 1]:
            simplex[n], fvals[n] = xr, fr
        else:
            shrink = True
            if fr < fvals[n]:
                if foc <= fr: simplex[n], fvals[n], shrink = xoc, foc, False
            else:
                if fic < fvals[n]: simplex[n], fvals[n], shrink = xic, fic, False
            if shrink:
                simplex[1:] = simplex[0] + delta * (simplex[1:] - simplex[0])
                fvals[1:] = f_batch(simplex[1:])
    best = int(np.argmin(fvals))
    return simplex[best], float(fvals[best]), converged


def optimize_gp_hyperparameters(y_obs_dim, t_obs, kernel_type: str, initial_log_params, jitter: float = 1e-6,
                                iterations: int = 100, g_tol: float = 1e-8, device: int = 0):
    """``optimize_gp_hyperparameters`` (src/initialization.jl:211-252): returns [variance, lengthscale, sigma]; falls back to the
    exponentiated initial guess when the optimum is not finite and positive (:240-247)."""
    x0 = np.asarray(initial_log_params, dtype=np.float64)
    f = lambda pts: negative_log_marginal_likelihood_batched(pts, y_obs_dim, t_obs, kernel_type, jitter, device)
    xbest, fbest, _ = _nelder_mead(f, x0, iterations=iterations, g_tol=g_tol)
    params = np.exp(xbest)
    if np.any(~np.isfinite(params)) or np.any(params <= 0) or not np.isfinite(fbest):
        return np.exp(x0)
    return params


def initial_guess(y_dim, t_obs):
    """Data-driven starting point of solve_magi (src/MagiJl.jl:277-294): [log var, log len, log sigma]."""
    y = np.asarray(y_dim, dtype=np.float64)
    t = np.asarray(t_obs, dtype=np.float64)
    v = y[~np.isnan(y)]
    time_range = float(t.max() - t.min())
    if v.size > 1:
        var_y = float(np.var(v, ddof=1))
        data_range = float(v.max() - v.min())
        mad = float(np.median(np.abs(v - np.median(v))) * 1.4826)
        return np.array([np.log(max(var_y, 1e-4)), np.log(max(time_range / 10.0, 1e-2)), np.log(max(mad, 1e-3 * data_range, 1e-4))])
    return np.array([np.log(1.0), np.log(max(time_range / 10.0, 1e-2)), np.log(0.1)])
