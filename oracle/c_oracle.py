"""ctypes binding for oracle/libmagi_oracle.so (the C restatement; test infrastructure / CPU baseline)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libmagi_oracle.so")
    src = os.path.join(_HERE, "magi_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmagi_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libmagi_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _tables(covs):
    m = np.ascontiguousarray(np.stack([np.asarray(g.mphiBand, dtype=np.float64) for g in covs]))
    k = np.ascontiguousarray(np.stack([np.asarray(g.KinvBand, dtype=np.float64) for g in covs]))
    c = np.ascontiguousarray(np.stack([np.asarray(g.CinvBand, dtype=np.float64) for g in covs]))
    return m, k, c


def batched(target, params, nthreads: int = 0):
    """params: (n_chains, P) C-contiguous.  Returns (ll[n_chains], grad[n_chains, P])."""
    params = np.ascontiguousarray(params, dtype=np.float64)
    nc, P = params.shape
    n, D, k = target.n_times, target.n_dims, target.n_params_ode
    m, kk, c = _tables(target.gp_cov_all_dims)
    Y = np.asfortranarray(np.asarray(target.yobs, dtype=np.float64))
    sig = np.ascontiguousarray(target.sigma_init, dtype=np.float64)
    beta = np.ascontiguousarray(target.prior_temperature, dtype=np.float64)
    ll = np.zeros(nc)
    grad = np.zeros((nc, P))
    b = int(target.gp_cov_all_dims[0].bandsize)
    rc = lib().magi_oracle_batched(n, D, k, target.model.model_id, b, int(target.sigma_is_fixed), nc, int(nthreads),
                                   _p(params), _p(sig), _p(Y.ravel(order="F").copy()), _p(m), _p(kk), _p(c), _p(beta), _p(ll), _p(grad))
    if rc != 0:
        raise RuntimeError("magi_oracle_batched failed")
    return ll, grad


def num_threads() -> int:
    return int(lib().magi_oracle_num_threads())
