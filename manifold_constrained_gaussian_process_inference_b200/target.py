"""MagiTarget and the LogDensityProblems-shaped interface (reference: src/logdensityproblems_interface.jl:33-45,
53-70, 111-166, 176-267).  Same names, argument meaning and error behaviour as the reference; the evaluation is one
call into libmagi_b200.so (fused sm_100a kernel) and there is no CPU path."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .gaussian_process import GPCov
from .ode_models import OdeSystem


class LogDensityOrder:
    """``LogDensityProblems.LogDensityOrder{K}``."""
    def __init__(self, order: int):
        self.order = order

    def __eq__(self, other):
        return isinstance(other, LogDensityOrder) and other.order == self.order

    def __repr__(self):
        return "LogDensityOrder{%d}()" % self.order


class MagiTarget:
    """``MagiTarget(yobs, gp_cov_all_dims, ode_system, sigma_init, prior_temperature, n_times, n_dims, n_params_ode,
    sigma_is_fixed)`` -- the reference's positional constructor (src/MagiJl.jl:508-520) with its three ODE callables
    replaced by an :class:`OdeSystem` naming a compiled device model.

    ``gp_cov_all_dims`` is a list of D :class:`GPCov`; their band tables (CinvBand, mphiBand, KinvBand) are uploaded
    to the GPU as they are, exactly as the reference reads them (src/likelihoods.jl:129-133,192)."""

    def __init__(self, yobs, gp_cov_all_dims, ode_system: OdeSystem, sigma_init, prior_temperature, n_times: int,
                 n_dims: int, n_params_ode: int, sigma_is_fixed: bool, device: int = 0, max_chains: int = 0):
        L = _lib.lib()
        self.yobs = np.asfortranarray(np.asarray(yobs, dtype=np.float64))
        if self.yobs.shape != (n_times, n_dims):
            raise ValueError("Dimensions of yobs %r do not match (n_times, n_dims) = (%d, %d)" % (self.yobs.shape, n_times, n_dims))
        if len(gp_cov_all_dims) != n_dims:
            raise ValueError("Length of gp_cov_all_dims (%d) does not match number of dimensions (%d)" % (len(gp_cov_all_dims), n_dims))
        self.gp_cov_all_dims = list(gp_cov_all_dims)
        self.ode_system = ode_system
        self.sigma_init = np.ascontiguousarray(sigma_init, dtype=np.float64)
        self.prior_temperature = np.ascontiguousarray(prior_temperature, dtype=np.float64)
        if self.prior_temperature.shape != (3,):
            raise ValueError("Length of prior_temperature must be 3.")
        if self.sigma_init.shape != (n_dims,):
            raise ValueError("Length of sigma does not match number of dimensions")
        self.n_times, self.n_dims, self.n_params_ode = int(n_times), int(n_dims), int(n_params_ode)
        self.sigma_is_fixed = bool(sigma_is_fixed)
        self.device = int(device)
        b = int(self.gp_cov_all_dims[0].bandsize)
        for g in self.gp_cov_all_dims:
            if int(g.bandsize) != b:
                raise ValueError("all dimensions must share one bandsize")
            for name in ("CinvBand", "mphiBand", "KinvBand"):
                T = getattr(g, name)
                if T is None or T.shape != (2 * b + 1, n_times):
                    raise ValueError("Pre-calculated GP covariance matrices have incorrect size (%s)" % name)
        self.bandsize = b
        tvec = np.ascontiguousarray(self.gp_cov_all_dims[0].tvec, dtype=np.float64)
        if tvec.shape[0] != n_times:
            raise ValueError("Length of tvec in GPCov does not match n_times")
        cfg = _lib.MagiConfig(
            n_times=self.n_times, n_dims=self.n_dims, n_params_ode=self.n_params_ode, kernel_id=0, bandsize=b,
            ode_model_id=ode_system.model_id, sigma_is_fixed=int(self.sigma_is_fixed), setup_mode=_lib.SETUP_INJECT,
            max_chains=int(max_chains), device=self.device, jitter=0.0, tvec=_lib.as_dp(tvec), phi=None,
            yobs=_lib.as_dp(self.yobs.ravel(order="F")), sigma_init=_lib.as_dp(self.sigma_init),
            prior_temperature=_lib.as_dp(self.prior_temperature))
        self._keep = (tvec,)
        h = ctypes.c_void_p()
        _lib.check(L.magi_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h
        self._L = L
        for d, g in enumerate(self.gp_cov_all_dims):
            for which, name in ((_lib.MAT_CINV_BAND, "CinvBand"), (_lib.MAT_MPHI_BAND, "mphiBand"), (_lib.MAT_KINV_BAND, "KinvBand")):
                T = np.ascontiguousarray(getattr(g, name), dtype=np.float64)
                _lib.check(L.magi_set_band_tables(h, d, which, _lib.as_dp(T)))

    @classmethod
    def from_config(cls, yobs, tvec, phi_all_dims, ode_system: OdeSystem, sigma_init, prior_temperature=(1.0, 1.0, 1.0),
                    sigma_is_fixed: bool = False, kernel: str = "matern52", bandsize: int = 20, jitter: float = 1e-6,
                    setup_mode: str = "reference_order", device: int = 0, max_chains: int = 0) -> "MagiTarget":
        """solve_magi's steps 4-5 in one call (src/MagiJl.jl:456-520): per-dimension GP setup from φ (2×D: row 0 variance,
        row 1 lengthscale, :466-467) and target construction, all on the GPU (K3-K6 run inside ``magi_create``); nothing but
        tvec, φ and yobs crosses PCIe.  bandsize is clamped to n−1 (:459)."""
        L = _lib.lib()
        self = cls.__new__(cls)
        self.yobs = np.asfortranarray(np.asarray(yobs, dtype=np.float64))
        n, D = self.yobs.shape
        phi = np.asarray(phi_all_dims, dtype=np.float64)
        if phi.shape != (2, D):
            raise ValueError("phi_all_dims must be 2 x D")
        self.gp_cov_all_dims = None
        self.ode_system = ode_system
        self.sigma_init = np.ascontiguousarray(sigma_init, dtype=np.float64)
        self.prior_temperature = np.ascontiguousarray(prior_temperature, dtype=np.float64)
        self.n_times, self.n_dims, self.n_params_ode = int(n), int(D), int(ode_system.thetaSize)
        self.sigma_is_fixed = bool(sigma_is_fixed)
        self.device = int(device)
        self.bandsize = max(0, min(int(bandsize), n - 1))
        tv = np.ascontiguousarray(tvec, dtype=np.float64)
        phi_flat = np.ascontiguousarray(phi.T.reshape(-1))            # [var_0, len_0, var_1, len_1, ...]
        cfg = _lib.MagiConfig(
            n_times=n, n_dims=D, n_params_ode=self.n_params_ode, kernel_id={"matern52": _lib.KERNEL_MATERN52, "rbf": _lib.KERNEL_RBF}[kernel],
            bandsize=int(bandsize), ode_model_id=ode_system.model_id, sigma_is_fixed=int(self.sigma_is_fixed),
            setup_mode={"reference_order": _lib.SETUP_REFERENCE_ORDER, "stable": _lib.SETUP_STABLE}[setup_mode],
            max_chains=int(max_chains), device=self.device, jitter=float(jitter), tvec=_lib.as_dp(tv), phi=_lib.as_dp(phi_flat),
            yobs=_lib.as_dp(self.yobs.ravel(order="F")), sigma_init=_lib.as_dp(self.sigma_init),
            prior_temperature=_lib.as_dp(self.prior_temperature))
        h = ctypes.c_void_p()
        _lib.check(L.magi_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h, self._L = h, L
        return self

    def get_matrix(self, dim: int, name: str) -> np.ndarray:
        """Dense GPCov field of one dimension (only after a device setup): C, Cinv, Cprime, Cdoubleprime, mphi, Kphi, Kinv."""
        which = {"C": _lib.MAT_C, "Cinv": _lib.MAT_CINV, "Cprime": _lib.MAT_CPRIME, "Cdoubleprime": _lib.MAT_CDOUBLEPRIME,
                 "mphi": _lib.MAT_MPHI, "Kphi": _lib.MAT_KPHI, "Kinv": _lib.MAT_KINV}[name]
        out = np.empty((self.n_times, self.n_times), order="F")
        _lib.check(self._L.magi_get_matrix(self._h, dim, which, _lib.as_dp(out)))
        return out

    def setup_status(self, dim: int):
        a, b = ctypes.c_int(), ctypes.c_int()
        _lib.check(self._L.magi_setup_status(self._h, dim, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)

    def setup_timing(self):
        """(kernel_ms, alloc_ms) of the device GP setup inside ``magi_create``: device time of K3-K6, host time in cudaMalloc."""
        k, a = ctypes.c_double(), ctypes.c_double()
        _lib.check(self._L.magi_setup_timing(self._h, ctypes.byref(k), ctypes.byref(a)))
        return float(k.value), float(a.value)

    # ---- LogDensityProblems interface ----
    def dimension(self) -> int:
        return int(self._L.magi_dimension(self._h))

    def capabilities(self) -> LogDensityOrder:
        return LogDensityOrder(int(self._L.magi_capabilities_order(self._h)))

    def logdensity(self, params) -> float:
        p = np.ascontiguousarray(params, dtype=np.float64)
        ll = ctypes.c_double()
        _lib.check(self._L.magi_logdensity(self._h, _lib.as_dp(p), int(p.shape[0]), ctypes.byref(ll)))
        return float(ll.value)

    def logdensity_and_gradient(self, params):
        p = np.ascontiguousarray(params, dtype=np.float64)
        P = self.dimension()
        grad = np.empty(P)
        ll = ctypes.c_double()
        _lib.check(self._L.magi_logdensity_and_gradient(self._h, _lib.as_dp(p), int(p.shape[0]), ctypes.byref(ll), _lib.as_dp(grad)))
        return float(ll.value), grad

    # ---- batched extensions (independent chains; the reference runs one chain, src/samplers.jl:173-184) ----
    def logdensity_and_gradient_batched(self, params, want_grad: bool = True):
        """params: (n_chains, P) array (row c = chain c, i.e. a Julia P×n_chains Matrix).  Returns (ll[n_chains], grad)."""
        p = np.ascontiguousarray(params, dtype=np.float64)
        if p.ndim != 2 or p.shape[1] != self.dimension():
            raise ValueError("params must be (n_chains, %d)" % self.dimension())
        nc = p.shape[0]
        ll = np.empty(nc)
        grad = np.empty_like(p) if want_grad else None
        _lib.check(self._L.magi_logdensity_and_gradient_batched(self._h, nc, _lib.as_dp(p), _lib.as_dp(ll), _lib.as_dp(grad) if want_grad else None))
        return ll, grad

    def logdensity_and_gradient_batched_dev(self, n_chains: int, params_ptr: int, ll_ptr: int, grad_ptr: int, stream: int = 0):
        """Device-resident evaluation: raw device pointers (e.g. torch ``tensor.data_ptr()``) and a CUDA stream handle."""
        _lib.check(self._L.magi_logdensity_and_gradient_batched_dev(self._h, int(n_chains), ctypes.c_void_p(params_ptr),
                                                                    ctypes.c_void_p(ll_ptr), ctypes.c_void_p(grad_ptr) if grad_ptr else None,
                                                                    _lib.LAYOUT_CHAIN_CONTIGUOUS, ctypes.c_void_p(stream)))

    def launch_count(self) -> int:
        return int(self._L.magi_launch_count(self._h))

    def get_band_table(self, dim: int, name: str) -> np.ndarray:
        which = {"CinvBand": _lib.MAT_CINV_BAND, "mphiBand": _lib.MAT_MPHI_BAND, "KinvBand": _lib.MAT_KINV_BAND}[name]
        out = np.empty((2 * self.bandsize + 1, self.n_times))
        _lib.check(self._L.magi_get_matrix(self._h, dim, which, _lib.as_dp(out)))
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.magi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# module-level functions with the reference's names (LogDensityProblems.dimension(target) etc.)
def dimension(target: MagiTarget) -> int:
    return target.dimension()


def capabilities(target) -> LogDensityOrder:
    return LogDensityOrder(1)


def logdensity(target: MagiTarget, params) -> float:
    return target.logdensity(params)


def logdensity_and_gradient(target: MagiTarget, params):
    return target.logdensity_and_gradient(params)
