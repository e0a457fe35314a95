"""ctypes binding of libmagi_b200.so (include/magi_b200.h).  There is no CPU fallback: if the library is missing
or no CUDA device is present, every compute entry point raises."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", os.environ.get("MAGI_LIB_NAME", "libmagi_b200.so"))

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)


class MagiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmagi_b200 error %d: %s" % (code, msg))
        self.code = code


class MagiConfig(ctypes.Structure):
    _fields_ = [
        ("n_times", ctypes.c_int), ("n_dims", ctypes.c_int), ("n_params_ode", ctypes.c_int), ("kernel_id", ctypes.c_int),
        ("bandsize", ctypes.c_int), ("ode_model_id", ctypes.c_int), ("sigma_is_fixed", ctypes.c_int), ("setup_mode", ctypes.c_int),
        ("max_chains", ctypes.c_int), ("device", ctypes.c_int), ("jitter", ctypes.c_double),
        ("tvec", c_double_p), ("phi", c_double_p), ("yobs", c_double_p), ("sigma_init", c_double_p), ("prior_temperature", c_double_p),
    ]


# status codes / enums of include/magi_b200.h
OK, ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_NOT_READY, ERR_UNSUPPORTED, ERR_NOT_POSITIVE_DEFINITE = range(6)
KERNEL_MATERN52, KERNEL_RBF, KERNEL_MATERN_NU12, KERNEL_MATERN_NU32, KERNEL_MATERN_NU52 = range(5)
SETUP_REFERENCE_ORDER, SETUP_STABLE, SETUP_INJECT = 0, 1, 2
MAT_C, MAT_CINV, MAT_CPRIME, MAT_CDOUBLEPRIME, MAT_MPHI, MAT_KPHI, MAT_KINV, MAT_CINV_BAND, MAT_MPHI_BAND, MAT_KINV_BAND = range(10)
LAYOUT_CHAIN_CONTIGUOUS = 0

_lib = None

# int (*magi_allreduce_fn)(void* dev_ptr, long long n_doubles, void* stream, void* user)  -- include/magi_b200.h
ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p)

# every symbol include/magi_b200.h declares (tests check the .so exports each of them)
EXPORTED = [
    "magi_last_error", "magi_version", "magi_create", "magi_destroy", "magi_dimension", "magi_capabilities_order",
    "magi_logdensity", "magi_logdensity_and_gradient", "magi_logdensity_and_gradient_batched",
    "magi_logdensity_and_gradient_batched_dev", "magi_get_matrix", "magi_set_band_tables", "magi_setup_status",
    "magi_launch_count", "magi_gp_covariances", "magi_gp_nlml_batched", "magi_hmc_init", "magi_hmc_run", "magi_hmc_reset_stats",
    "magi_hmc_get_state", "magi_hmc_get_draws", "magi_hmc_draws_dev", "magi_hmc_get_stats", "magi_hmc_grad_evals",
    "magi_hmc_set_global", "magi_setup_timing", "magi_nccl_unique_id", "magi_comm_init", "magi_comm_attach", "magi_comm_warmup",
    "magi_hmc_allgather_draws", "magi_hmc_store_x", "magi_hmc_get_x_draws", "magi_nuts_run", "magi_nuts_get_stats",
]


def lib():
    """Loads the shared library (raises if it has not been built: run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MagiError(-1, "%s not found: build it with `python -m manifold_constrained_gaussian_process_inference_b200.build` "
                            "(there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, dp = ctypes.c_void_p, ctypes.c_int, c_double_p
    L.magi_last_error.restype = ctypes.c_char_p
    L.magi_create.argtypes = [ctypes.POINTER(MagiConfig), ctypes.POINTER(vp)]
    L.magi_destroy.argtypes = [vp]
    L.magi_dimension.argtypes = [vp]
    L.magi_capabilities_order.argtypes = [vp]
    L.magi_logdensity.argtypes = [vp, dp, ci, dp]
    L.magi_logdensity_and_gradient.argtypes = [vp, dp, ci, dp, dp]
    L.magi_logdensity_and_gradient_batched.argtypes = [vp, ci, dp, dp, dp]
    L.magi_logdensity_and_gradient_batched_dev.argtypes = [vp, ci, vp, vp, vp, ci, vp]
    L.magi_get_matrix.argtypes = [vp, ci, ci, dp]
    L.magi_set_band_tables.argtypes = [vp, ci, ci, dp]
    L.magi_setup_status.argtypes = [vp, ci, c_int_p, c_int_p]
    L.magi_setup_timing.argtypes = [vp, dp, dp]
    L.magi_launch_count.argtypes = [vp]
    L.magi_launch_count.restype = ctypes.c_longlong
    for name in dir(L):
        pass
    _optional(L)
    _lib = L
    return L


def _optional(L):
    """Entry points added after the first milestone; bound when present."""
    vp, ci, dp = ctypes.c_void_p, ctypes.c_int, c_double_p
    if hasattr(L, "magi_gp_covariances"):
        L.magi_gp_covariances.argtypes = [ci, dp, dp, ci, ci, ctypes.c_double, ci, ci, ci, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, c_int_p]
    if hasattr(L, "magi_gp_nlml_batched"):
        L.magi_gp_nlml_batched.argtypes = [ci, ci, dp, dp, ctypes.c_double, ci, dp, dp, ci]
    if hasattr(L, "magi_hmc_init"):
        ll = ctypes.c_longlong
        L.magi_hmc_init.argtypes = [vp, ci, dp, ctypes.c_ulonglong, ctypes.c_double, ll]
        L.magi_hmc_run.argtypes = [vp, ci, ci, ci, ctypes.c_double, ci, vp]
        L.magi_hmc_reset_stats.argtypes = [vp]
        L.magi_hmc_get_state.argtypes = [vp, dp, dp]
        L.magi_hmc_get_draws.argtypes = [vp, dp, ll, ctypes.POINTER(ll)]
        L.magi_hmc_draws_dev.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(ll), c_int_p, c_int_p]
        L.magi_hmc_get_stats.argtypes = [vp, dp, dp, c_int_p, dp, dp]
        L.magi_nuts_run.argtypes = [vp, ci, ci, ci, ctypes.c_double, ci, vp]
        L.magi_nuts_get_stats.argtypes = [vp, dp, dp]
        L.magi_hmc_store_x.argtypes = [vp, ci, ci]
        L.magi_hmc_get_x_draws.argtypes = [vp, dp, ll, ctypes.POINTER(ll), c_int_p]
        L.magi_hmc_grad_evals.argtypes = [vp]
        L.magi_hmc_grad_evals.restype = ll
    if hasattr(L, "magi_comm_init"):
        L.magi_nccl_unique_id.argtypes = [ctypes.c_char_p]
        L.magi_comm_init.argtypes = [vp, ctypes.c_char_p, ci, ci]
        L.magi_comm_attach.argtypes = [vp, vp, ci, ci]
        L.magi_comm_warmup.argtypes = [vp, vp]
        L.magi_hmc_allgather_draws.argtypes = [vp, vp, vp]
    if hasattr(L, "magi_hmc_set_global"):
        L.magi_hmc_set_global.argtypes = [vp, ctypes.c_longlong, ALLREDUCE_FN, vp]


def check(rc):
    if rc != 0:
        raise MagiError(rc, lib().magi_last_error().decode("utf-8", "replace"))


def as_dp(a):
    return a.ctypes.data_as(c_double_p)
