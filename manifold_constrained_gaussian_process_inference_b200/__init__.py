"""B200-native MAGI hot path: log-posterior + gradient (and the GP setup feeding it) behind the reference's
LogDensityProblems boundary.  The compute lives in csrc/ (hand-written sm_100a CUDA behind the C ABI of
include/magi_b200.h); this package is the host-side mirror of the reference interface
(MagiJl.jl: GPCov / calculate_gp_covariances!, MagiTarget, dimension / logdensity / logdensity_and_gradient,
run_nuts_sampler-shaped batched HMC, solve_magi)."""
from . import _lib
from .kernels import Kernel, create_general_matern_kernel, create_matern52_kernel, create_rbf_kernel
from .ode_models import OdeSystem, get_ode_system, fn_system, hes1_system, lv_system, MODEL_IDS
from .gaussian_process import GPCov, calculate_gp_covariances, mat2band
from .samplers import run_hmc_sampler, run_nuts_sampler, logdensity_func_wrapper, logdensity_and_gradient_func_wrapper
from .solver import solve_magi
from . import diagnostics, distributed, initialization
from .target import MagiTarget, dimension, capabilities, logdensity, logdensity_and_gradient, LogDensityOrder

__all__ = [
    "Kernel", "create_matern52_kernel", "create_rbf_kernel", "create_general_matern_kernel", "OdeSystem", "get_ode_system", "fn_system", "hes1_system",
    "lv_system", "MODEL_IDS", "GPCov", "calculate_gp_covariances", "mat2band", "MagiTarget", "dimension", "capabilities",
    "logdensity", "logdensity_and_gradient", "LogDensityOrder", "run_hmc_sampler", "run_nuts_sampler", "logdensity_func_wrapper", "logdensity_and_gradient_func_wrapper", "solve_magi", "diagnostics", "distributed",
]
