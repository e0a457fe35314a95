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
                    n_chains_total: int | None = None, window_allreduce=None, stream: int = 0):
    """Argument meaning follows ``run_nuts_sampler``: ``n_samples`` is the TOTAL number of iterations including the
    ``n_adapts`` warm-up iterations, which are dropped (``drop_warmup=true``).  ``initial_params`` is (n_chains, P).

    Multi-rank runs (chains sharded over GPUs): call ``distributed.init_device_comm(target)`` first and pass the global chain
    count as ``n_chains_total``; the pooled metric of the warm-up is then a statistic of ALL ranks' chains (one ncclAllReduce
    per adaptation window inside the library), and the run is bit-identical to a one-rank run over the same global chains
    (shards aligned to ``n_chains_total / 64`` chains).  ``window_allreduce`` (``distributed.make_window_allreduce``) is the
    host-callback alternative for transports other than NCCL.  ``stream``: CUDA stream handle the sampler runs on (0: the
    handle's own stream).

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
    if n_adapts > 0:
        _lib.check(L.magi_hmc_run(h, int(n_adapts), int(n_leapfrog), 1, float(target_accept_ratio), 0, ctypes.c_void_p(stream) if stream else None))
    _lib.check(L.magi_hmc_reset_stats(h))
    n_keep = int(n_samples) - int(n_adapts)
    if n_keep > 0:
        _lib.check(L.magi_hmc_run(h, n_keep, int(n_leapfrog), 0, float(target_accept_ratio), 1, ctypes.c_void_p(stream) if stream else None))
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
    return chain, stats


def hmc_draws_device_view(target: MagiTarget):
    """(device pointer, n_stored, n_chains, n_cols) of the on-device draw store, for a NCCL all-gather without a host copy."""
    L = _lib.lib()
    ptr = ctypes.c_void_p(); ns = ctypes.c_longlong(); nc = ctypes.c_int(); ncol = ctypes.c_int()
    _lib.check(L.magi_hmc_draws_dev(target._h, ctypes.byref(ptr), ctypes.byref(ns), ctypes.byref(nc), ctypes.byref(ncol)))
    return ptr.value, int(ns.value), int(nc.value), int(ncol.value)
