"""solve_magi entry (reference: src/MagiJl.jl:170-773) with the hot path on the GPU.

Kept from the reference: the config keys and defaults (:kernel "matern52", :niterHmc 20000, :burninRatio 0.5,
:stepSizeFactor 0.01, :bandSize 20, :priorTemperature [1,1,1], :jitter 1e-6, :targetAcceptRatio 0.8, :sigma, :phi, :xInit,
:thetaInit; :208-220), ``sigma_is_fixed = :sigma and :phi both given`` (:224), linear-interpolation X init (:351-410),
bounds-based θ init (:412-453), band clamp (:459), parameter vector layout [vec(X); θ; log σ] (:526-569), burn-in split
(:578-581) and the shape of the result (θ, x_sampled, σ, φ, lp; :633-771).  New keys: :nChains (independent chains, default 1024),
:nLeapfrog (static trajectory length), :maxTreeDepth (> 0: batched NUTS trees instead, the reference's sampler), :setupMode, :seed, :device, :xChains (chains whose latent trajectories are kept, default
min(nChains, 16)), :xThin (keep X at every xThin-th kept iteration, default 1).

When ``config['phi']`` and/or ``config['sigma']`` are absent they are estimated per dimension by minimising the GP negative log
marginal likelihood (src/MagiJl.jl:254-330 -> src/initialization.jl), objective on the GPU, Nelder-Mead on the host
(``initialization.py``); ``config['sigmaInit']`` overrides the estimated starting σ."""
from __future__ import annotations

import numpy as np

from . import initialization
from .ode_models import OdeSystem
from .samplers import run_hmc_sampler
from .target import MagiTarget


def initial_x(y_obs, t_obs):
    """Linear interpolation / extrapolation of the finite observations per dimension (src/MagiJl.jl:351-401)."""
    y = np.asarray(y_obs, dtype=np.float64)
    t = np.asarray(t_obs, dtype=np.float64)
    n, D = y.shape
    x = np.zeros((n, D))
    for d in range(D):
        ok = np.where(~np.isnan(y[:, d]))[0]
        if len(ok) == 0:
            continue
        if len(ok) < 2:
            x[:, d] = y[ok[0], d]
            continue
        tt, idx = np.unique(t[ok], return_index=True)
        vv = y[ok[idx], d]
        if len(tt) < 2:
            x[:, d] = vv[0]
            continue
        x[:, d] = np.interp(t, tt, vv)
        lo, hi = t < tt[0], t > tt[-1]                      # extrapolation_bc = Line()
        x[lo, d] = vv[0] + (t[lo] - tt[0]) * (vv[1] - vv[0]) / (tt[1] - tt[0])
        x[hi, d] = vv[-1] + (t[hi] - tt[-1]) * (vv[-1] - vv[-2]) / (tt[-1] - tt[-2])
    return x


def initial_theta(ode_system: OdeSystem):
    """Bounds-based θ₀ (src/MagiJl.jl:412-439)."""
    k = ode_system.thetaSize
    th = np.zeros(k)
    for i in range(k):
        lb, ub = ode_system.thetaLowerBound[i], ode_system.thetaUpperBound[i]
        if np.isfinite(lb) and np.isfinite(ub):
            th[i] = (lb + ub) / 2.0
        elif np.isfinite(lb):
            th[i] = lb + abs(lb) * 0.1 + 0.1
        elif np.isfinite(ub):
            th[i] = ub - abs(ub) * 0.1 - 0.1
        if np.isfinite(lb) and th[i] <= lb:
            th[i] = lb + 1e-4 * (min(1.0, ub - lb) if np.isfinite(ub) else 1.0)
        if np.isfinite(ub) and th[i] >= ub:
            th[i] = ub - 1e-4 * (min(1.0, ub - lb) if np.isfinite(lb) else 1.0)
        th[i] = min(max(th[i], lb), ub)
    return th


def solve_magi(y_obs, t_obs, ode_system: OdeSystem, config=None, initial_params=None):
    cfg = dict(config or {})
    get = lambda k, d: cfg.get(k, d)
    y = np.asarray(y_obs, dtype=np.float64)
    t = np.asarray(t_obs, dtype=np.float64)
    n, D = y.shape
    k = ode_system.thetaSize
    kernel = get("kernel", "matern52")
    niter = int(get("niterHmc", 20000))
    burn = float(get("burninRatio", 0.5))
    eps0 = float(get("stepSizeFactor", 0.01))
    band = int(get("bandSize", 20))
    beta = list(get("priorTemperature", [1.0, 1.0, 1.0]))
    if len(beta) != 3:
        beta = [float(beta[0])] * 3                                        # :498-501
    jitter = float(get("jitter", 1e-6))
    delta = float(get("targetAcceptRatio", 0.8))
    phi = get("phi", None)
    sigma = get("sigma", None)
    sigma_is_fixed = (sigma is not None) and (phi is not None)               # :224
    if sigma is not None and phi is None:                                   # :233-236: sigma without phi is discarded and re-initialised
        import warnings
        warnings.warn("Sigma provided but Phi not provided. Sigma will be treated as unknown and re-initialized.")
        sigma = None
    device = int(get("device", 0))
    sigma_est = None
    if phi is None or sigma is None:                                        # :254-330: Option A, estimate per dimension
        phi_est = np.zeros((2, D))
        sigma_est = np.zeros(D)
        for d in range(D):
            x0 = initialization.initial_guess(y[:, d], t)
            if phi is not None:
                x0[:2] = np.log(np.asarray(phi, dtype=np.float64).reshape(2, D)[:, d])
            try:
                opt = initialization.optimize_gp_hyperparameters(y[:, d], t, kernel, x0, jitter=jitter,
                                                                 iterations=int(get("gpOptimIterations", 100)),
                                                                 g_tol=float(get("gpOptimGTol", 1e-8)), device=device)
            except Exception:                                               # :313-317 fall back to the initial guess
                opt = np.exp(x0)
            phi_est[:, d] = opt[:2]
            sigma_est[d] = max(opt[2], 1e-8)                                # :326
        if phi is None:
            phi = phi_est
    phi = np.asarray(phi, dtype=np.float64).reshape(2, D)
    if np.any(~np.isfinite(phi)) or np.any(phi <= 0):
        raise ValueError("Invalid GP hyperparameters: variance and lengthscale must be finite and > 0")   # :469-472
    if sigma is not None:
        sigma_init = np.asarray(sigma, dtype=np.float64)
    elif get("sigmaInit", None) is not None:
        sigma_init = np.asarray(get("sigmaInit", None), dtype=np.float64)
    else:
        sigma_init = sigma_est
    n_chains = int(get("nChains", 1024))
    n_leap = int(get("nLeapfrog", 20))
    if kernel not in ("matern52", "rbf"):
        kernel = "matern52"                                                # :477-480
    x_init = np.asarray(get("xInit", None) if get("xInit", None) is not None else initial_x(y, t), dtype=np.float64)
    if x_init.shape != (n, D):
        raise ValueError("Provided :xInit matrix has wrong dimensions")
    th_init = get("thetaInit", None)
    th_init = initial_theta(ode_system) if th_init is None else np.clip(np.asarray(th_init, dtype=np.float64), ode_system.thetaLowerBound, ode_system.thetaUpperBound)
    target = MagiTarget.from_config(y, t, phi, ode_system, sigma_init, prior_temperature=beta, sigma_is_fixed=sigma_is_fixed,
                                    kernel=kernel, bandsize=band, jitter=jitter, setup_mode=get("setupMode", "reference_order"),
                                    device=device, max_chains=n_chains)
    P = target.dimension()
    if initial_params is None:
        parts = [x_init.reshape(-1, order="F"), th_init]
        if not sigma_is_fixed:
            parts.append(np.log(np.maximum(sigma_init, 1e-8)))             # :530
        p0 = np.concatenate(parts)
    else:
        p0 = np.asarray(initial_params, dtype=np.float64)
        if p0.shape[-1] != P:
            raise ValueError("Provided initial_params vector has wrong length. Expected %d" % P)
    if p0.ndim == 1:
        # over-dispersed starts around the reference's single starting point (all chains share the mode-finding start)
        rng = np.random.default_rng(int(get("seed", 0)))
        p0 = p0[None, :] + 0.01 * rng.normal(size=(n_chains, P)) * np.maximum(1.0, np.abs(p0))[None, :] * (np.arange(n_chains) > 0)[:, None]
    n_adapts = int(np.floor(niter * burn))                                 # :578
    x_chains = int(get("xChains", min(n_chains, 16)))
    chain, stats = run_hmc_sampler(target, p0, n_samples=niter, n_adapts=n_adapts, target_accept_ratio=delta,
                                   initial_step_size=eps0, n_leapfrog=n_leap, seed=int(get("seed", 0)),
                                   x_chains=x_chains, x_thin=int(get("xThin", 1)), max_tree_depth=int(get("maxTreeDepth", 0)))
    theta = chain[:, :, :k]
    sig = chain[:, :, k:k + D]
    if sigma_is_fixed:
        sig = np.broadcast_to(sigma_init[None, None, :], sig.shape).copy() # repeated rows (:696)
    lp = chain[:, :, k + D]
    x_sampled = stats.get("x_sampled")                                      # (S', xChains, n, D); S' = S when xThin == 1
    if n_chains == 1:                                                       # the reference's shapes: S×k, S×n×D, S×D, S (:633-771)
        theta, sig, lp = theta[:, 0], sig[:, 0], lp[:, 0]
        x_sampled = x_sampled[:, 0] if x_sampled is not None else None
    return dict(theta=theta, x_sampled=x_sampled, sigma=sig, phi=phi, lp=lp,
                x_mean=stats["x_mean"].reshape(n_chains, D, n).transpose(0, 2, 1), stats=stats, target=target)
