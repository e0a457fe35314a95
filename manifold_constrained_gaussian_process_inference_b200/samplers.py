"""Batched sampler entry (reference: run_nuts_sampler, src/samplers.jl:114-194).  The reference drives one NUTS chain
from the host; here thousands of independent chains run on-device static-trajectory HMC with the same adaptation
recipe (diagonal metric, dual-averaging step size towards ``target_accept_ratio``), every leapfrog step being one call
of the fused log-posterior/gradient kernel.  The drop-in single-chain boundary for the reference's own NUTS loop is
``MagiTarget.logdensity_and_gradient``."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .target import MagiTarget


def run_hmc_sampler(target: MagiTarget, initial_params, n_samples: int = 2000, n_adapts: int = 1000,
                    target_accept_ratio: float = 0.8, initial_step_size: float = 0.1, n_leapfrog: int = 20,
                    seed: int = 0, chain_id_offset: int = 0, keep_on_device: bool = False,
                    n_chains_total: int | None = None, window_allreduce=None, stream: int = 0,
                    x_chains: int = 0, x_thin: int = 1, max_tree_depth: int = 0):
    """Argument meaning follows ``run_nuts_sampler``: ``n_samples`` is the TOTAL number of iterations including the
    ``n_adapts`` warm-up iterations, which are dropped (``drop_warmup=true``).  ``initial_params`` is (n_chains, P).

    Multi-rank runs (chains sharded over GPUs): call ``distributed.init_device_comm(target)`` first and pass the global chain
    count as ``n_chains_total``; the pooled metric of the warm-up is then a statistic of ALL ranks' chains (one ncclAllReduce
    per adaptation window inside the library), and the run is bit-identical to a one-rank run over the same global chains
    (shards aligned to ``n_chains_total / 64`` chains).  ``window_allreduce`` (``distributed.make_window_allreduce``) is the
    host-callback alternative for transports other than NCCL.  ``stream``: CUDA stream handle the sampler runs on (0: the
    handle's own stream).

    ``max_tree_depth`` > 0 replaces the static trajectories by batched NUTS trees (multinomial sampling, generalised U-turn
    criterion: the reference's ``Trajectory{MultinomialTS}(Leapfrog, GeneralisedNoUTurn)``, src/samplers.jl:158-160) of at most
    ``2**max_tree_depth - 1`` leapfrog steps; ``n_leapfrog`` is then ignored and ``stats`` gains ``tree_depth`` / ``n_leapfrog_mean``.

    ``x_chains`` > 0 keeps the latent trajectories X of the first ``x_chains`` chains at every ``x_thin``-th kept iteration
    (``stats["x_sampled"]``, shape (n_stored, x_chains, n, D): the reference's ``x_sampled`` S×n×D per chain).

    Returns ``(chain, stats)``: ``chain`` is an array (n_kept, n_chains, k + D + 1) of (θ, σ, lp) draws and ``stats`` a
    dict with per-chain acceptance rate, step size, divergences, posterior mean of X, the adapted inverse metric and
    the number of gradient evaluations."""
    L = _lib.lib()
    p0 = np.ascontiguousarray(initial_params, dtype=np.float64)
    if p0.ndim == 1:
        p0 = p0[None, :]
    P = target.dimension()
    assert p0.shape[1] == P, "Initial parameters dimension mismatch"          # samplers.jl:125
    nc = p0.shape[0]
    h = target._h
    _lib.check(L.magi_hmc_init(h, nc, _lib.as_dp(p0), ctypes.c_ulonglong(seed), float(initial_step_size), ctypes.c_longlong(chain_id_offset)))
    if n_chains_total is not None or window_allreduce is not None:
        cb = window_allreduce if window_allreduce is not None else ctypes.cast(None, _lib.ALLREDUCE_FN)
        target._window_allreduce = cb                      # keep the ctypes callback alive as long as the handle
        _lib.check(L.magi_hmc_set_global(h, ctypes.c_longlong(int(n_chains_total if n_chains_total is not None else nc + chain_id_offset)), cb, None))
    stp = ctypes.c_void_p(stream) if stream else None
    if max_tree_depth > 0:
        run = lambda n_it, adapt, store: L.magi_nuts_run(h, int(n_it), int(max_tree_depth), adapt, float(target_accept_ratio), store, stp)
    else:
        run = lambda n_it, adapt, store: L.magi_hmc_run(h, int(n_it), int(n_leapfrog), adapt, float(target_accept_ratio), store, stp)
    if n_adapts > 0:
        _lib.check(run(n_adapts, 1, 0))
    _lib.check(L.magi_hmc_reset_stats(h))
    if x_chains > 0:
        _lib.check(L.magi_hmc_store_x(h, int(x_chains), int(x_thin)))
    n_keep = int(n_samples) - int(n_adapts)
    if n_keep > 0:
        _lib.check(run(n_keep, 0, 1))
    ncols = target.n_params_ode + target.n_dims + 1
    chain = None
    if not keep_on_device:
        chain = np.empty((max(n_keep, 0), nc, ncols))
        ns = ctypes.c_longlong()
        _lib.check(L.magi_hmc_get_draws(h, _lib.as_dp(chain), ctypes.c_longlong(max(n_keep, 0)), ctypes.byref(ns)))
    acc = np.empty(nc); eps = np.empty(nc); ndiv = np.empty(nc, dtype=np.int32)
    xmean = np.empty((nc, target.n_times * target.n_dims)); minv = np.empty(P)
    _lib.check(L.magi_hmc_get_stats(h, _lib.as_dp(acc), _lib.as_dp(eps), ndiv.ctypes.data_as(_lib.c_int_p), _lib.as_dp(xmean), _lib.as_dp(minv)))
    stats = dict(accept_rate=acc, step_size=eps, n_divergent=ndiv, x_mean=xmean, inverse_metric=minv,
                 grad_evals=int(L.magi_hmc_grad_evals(h)), n_leapfrog=int(n_leapfrog))
    if max_tree_depth > 0:
        td, nl = np.empty(nc), np.empty(nc)
        _lib.check(L.magi_nuts_get_stats(h, _lib.as_dp(td), _lib.as_dp(nl)))
        stats["tree_depth"], stats["n_leapfrog_mean"] = td, nl
    if x_chains > 0:
        ns, ncx = ctypes.c_longlong(), ctypes.c_int()
        _lib.check(L.magi_hmc_get_x_draws(h, None, ctypes.c_longlong(0), ctypes.byref(ns), ctypes.byref(ncx)))
        xs = np.empty((int(ns.value), int(ncx.value), target.n_dims, target.n_times))
        if xs.size:
            _lib.check(L.magi_hmc_get_x_draws(h, _lib.as_dp(xs), ns, ctypes.byref(ns), ctypes.byref(ncx)))
        stats["x_sampled"] = np.ascontiguousarray(xs.transpose(0, 1, 3, 2))       # (S, chains, n, D): vec(X) is time-fastest

    return chain, stats


def hmc_draws_device_view(target: MagiTarget):
    """(device pointer, n_stored, n_chains, n_cols) of the on-device draw store, for a NCCL all-gather without a host copy."""
    L = _lib.lib()
    ptr = ctypes.c_void_p(); ns = ctypes.c_longlong(); nc = ctypes.c_int(); ncol = ctypes.c_int()
    _lib.check(L.magi_hmc_draws_dev(target._h, ctypes.byref(ptr), ctypes.byref(ns), ctypes.byref(nc), ctypes.byref(ncol)))
    return ptr.value, int(ns.value), int(nc.value), int(ncol.value)


# ---------------------------------------------------------------------------------------------------------------------------
# The reference's own caller of the single-chain boundary: run_nuts_sampler (src/samplers.jl:114-194) drives
# LogDensityProblems.logdensity_and_gradient once per leapfrog step from ONE host task, through two wrappers that assert a
# finite value and gradient (:29-32, :53-63).  The reference delegates the NUTS transition to AdvancedHMC 0.7.1
# (Trajectory{MultinomialTS}(Leapfrog, GeneralisedNoUTurn), DiagEuclideanMetric, StanHMCAdaptor); that package is not under
# the reference tree, so the transition below is restated from its published algorithm (Hoffman & Gelman 2014; Betancourt
# 2017: multinomial sampling, generalised U-turn criterion on rho = sum of momenta; Stan's dual averaging and windowed
# diagonal metric) -- host logic, one chain, every gradient a call into magi_logdensity_and_gradient.
# ---------------------------------------------------------------------------------------------------------------------------

def logdensity_func_wrapper(target, theta):
    """``logdensity_func_wrapper`` (src/samplers.jl:29-32)."""
    return target.logdensity(theta)


def logdensity_and_gradient_func_wrapper(target, theta):
    """``logdensity_and_gradient_func_wrapper`` (src/samplers.jl:53-63): the assertions are the reference's."""
    val, grad = target.logdensity_and_gradient(theta)
    assert np.isreal(val) and np.isfinite(val), "Log density is not a finite Real! Value: %r" % (val,)
    assert isinstance(grad, np.ndarray) and grad.ndim == 1, "Gradient is not a Vector! Type: %s" % type(grad)
    assert np.all(np.isfinite(grad)), "Gradient contains non-finite values!"
    return val, grad


class _DualAveraging:
    """Nesterov dual averaging of log eps (Stan / AdvancedHMC defaults: gamma 0.05, t0 10, kappa 0.75, mu = log(10 eps0))."""
    def __init__(self, eps0, delta):
        self.delta = delta
        self.restart(eps0)

    def restart(self, eps0):
        self.mu, self.m, self.hbar, self.log_eps_bar, self.eps = np.log(10.0 * eps0), 0, 0.0, 0.0, eps0

    def update(self, accept_stat):
        self.m += 1
        eta = 1.0 / (self.m + 10.0)
        self.hbar = (1.0 - eta) * self.hbar + eta * (self.delta - accept_stat)
        log_eps = self.mu - np.sqrt(self.m) / 0.05 * self.hbar
        w = self.m ** -0.75
        self.log_eps_bar = w * log_eps + (1.0 - w) * self.log_eps_bar
        self.eps = float(np.exp(log_eps))

    def final(self):
        return float(np.exp(self.log_eps_bar))


def _leapfrog(lpg, q, p, g, eps, minv):
    p = p + 0.5 * eps * g
    q = q + eps * minv * p
    lp, g = lpg(q)
    p = p + 0.5 * eps * g
    return q, p, lp, g


def _nuts_transition(rng, lpg, q0, lp0, g0, eps, minv, max_depth=10, max_delta_h=1000.0):
    """One NUTS transition (multinomial sampling, generalised no-U-turn).  Returns (q, lp, g, stats)."""
    p0 = rng.normal(size=q0.shape) / np.sqrt(minv)
    h0 = -lp0 + 0.5 * float(p0 @ (minv * p0))
    stats = dict(n_leapfrog=0, sum_accept=0.0, divergent=False, depth=0)

    def uturn(rho, p_left, p_right):
        return float(rho @ (minv * p_left)) <= 0.0 or float(rho @ (minv * p_right)) <= 0.0

    def build(q, p, g, direction, depth):
        # returns (outer q, p, g), (inner-side p of the subtree: first state built), sample (q, lp, g), log weight, rho, valid
        if depth == 0:
            q1, p1, lp1, g1 = _leapfrog(lpg, q, p, g, direction * eps, minv)
            h1 = -lp1 + 0.5 * float(p1 @ (minv * p1))
            dh = h1 - h0
            if not np.isfinite(dh):
                dh = np.inf
            stats["n_leapfrog"] += 1
            stats["sum_accept"] += 1.0 if dh <= 0.0 else (float(np.exp(-dh)) if dh < np.inf else 0.0)
            div = dh > max_delta_h
            stats["divergent"] |= bool(div)
            return (q1, p1, g1), p1, (q1, lp1, g1), -dh, p1.copy(), not div
        o1, pin1, s1, w1, rho1, ok1 = build(q, p, g, direction, depth - 1)
        if not ok1:
            return o1, pin1, s1, w1, rho1, False
        o2, pin2, s2, w2, rho2, ok2 = build(o1[0], o1[1], o1[2], direction, depth - 1)
        w = np.logaddexp(w1, w2)
        s = s2 if (ok2 and np.log(rng.uniform()) < w2 - w) else s1          # uniform (multinomial) sampling inside a subtree
        rho = rho1 + rho2
        ok = ok2 and not uturn(rho, pin1, o2[1])
        return o2, pin1, s, w, rho, ok

    left = right = (q0, p0, g0)
    sample, logw, rho = (q0, lp0, g0), 0.0, p0.copy()
    for depth in range(max_depth):
        direction = 1 if rng.uniform() < 0.5 else -1
        start = right if direction == 1 else left
        outer, _, s_new, w_new, rho_new, ok = build(start[0], start[1], start[2], direction, depth)
        if not ok:
            break
        stats["depth"] = depth + 1
        if np.log(rng.uniform()) < w_new - logw:                             # biased progressive sampling between the old tree and the new subtree
            sample = s_new
        logw = np.logaddexp(logw, w_new)
        rho = rho + rho_new
        if direction == 1:
            right = outer
        else:
            left = outer
        if uturn(rho, left[1], right[1]):
            break
    stats["accept_stat"] = stats["sum_accept"] / max(1, stats["n_leapfrog"])
    return sample[0], sample[1], sample[2], stats


def run_nuts_sampler(target: MagiTarget, initial_params, n_samples: int = 20000, n_adapts: int = 10000,
                     target_accept_ratio: float = 0.8, initial_step_size: float = 0.1, seed: int = 0, max_depth: int = 10):
    """``run_nuts_sampler`` (src/samplers.jl:114-194): ONE chain, NUTS on the host, every gradient a call of the single-chain
    drop-in ``magi_logdensity_and_gradient``.  Same arguments and failure behaviour as the reference: the length of
    ``initial_params`` is asserted against ``dimension(target)`` (:125); a non-finite value or gradient anywhere raises inside
    the wrapper and the run returns ``(None, None)`` (:186-190).  Returns ``(chain, stats)``: ``chain`` (n_samples - n_adapts,
    P) with the warm-up dropped, ``stats`` a list of per-iteration dicts (n_leapfrog, depth, accept_stat, divergent, step_size)."""
    n_dims_total = target.dimension()
    theta0 = np.ascontiguousarray(initial_params, dtype=np.float64)
    assert theta0.ndim == 1 and theta0.shape[0] == n_dims_total, "Initial parameters dimension mismatch"
    try:
        lpg = lambda th: logdensity_and_gradient_func_wrapper(target, th)
        rng = np.random.default_rng(seed)
        minv = np.ones(n_dims_total)                                         # DiagEuclideanMetric: M^-1 = I
        da = _DualAveraging(float(initial_step_size), float(target_accept_ratio))
        q = theta0.copy()
        lp, g = lpg(q)
        # Stan-style windows over the warm-up: initial fast buffer, doubling slow windows (metric), terminal fast buffer
        init_buf, term_buf, base_win = 75, 50, 25
        if n_adapts < 150:
            init_buf, term_buf = int(0.15 * n_adapts), int(0.10 * n_adapts)
            base_win = n_adapts - init_buf - term_buf
        win_end, win_size, win = init_buf + base_win, base_win, []
        if n_adapts - term_buf - win_end < 2 * win_size:
            win_end = n_adapts - term_buf
        chain, stats = [], []
        eps_final = float(initial_step_size)
        for it in range(int(n_samples)):
            q, lp, g, st = _nuts_transition(rng, lpg, q, lp, g, da.eps if it < n_adapts else eps_final, minv, max_depth)
            st["step_size"] = da.eps if it < n_adapts else eps_final
            if it < n_adapts:
                da.update(st["accept_stat"])
                if init_buf <= it < n_adapts - term_buf:
                    win.append(q.copy())
                    if it + 1 == win_end and len(win) > 1:
                        w = np.asarray(win)
                        nw = w.shape[0]
                        minv = w.var(axis=0, ddof=1) * (nw / (nw + 5.0)) + 1e-3 * (5.0 / (nw + 5.0))   # Stan's regularised variance
                        win, win_size = [], win_size * 2
                        nxt = win_end + win_size
                        win_end = n_adapts - term_buf if n_adapts - term_buf - nxt < 2 * win_size else nxt
                        da.restart(da.eps)
                if it + 1 == n_adapts:
                    eps_final = da.final()
            else:
                chain.append(q.copy())
                stats.append(st)
        return np.asarray(chain), stats
    except AssertionError as e:                                               # samplers.jl:186-190
        import warnings
        warnings.warn("ERROR in NUTS sampler! %s" % (e,))
        return None, None
