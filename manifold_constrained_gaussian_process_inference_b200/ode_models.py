"""ODE model registry (reference: src/ode_models.jl).  The reference passes Julia closures (f!, dfdx!, dfdp) in an
``OdeSystem``; closures cannot cross a C ABI into a kernel, so an OdeSystem here names a compiled device model."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

MODEL_IDS = {"fn": 0, "hes1": 1, "hes1log": 2, "hes1log_fixg": 3, "hes1log_fixf": 4, "hiv": 5, "ptrans": 6, "lv": 7, "lorenz96": 8}
_DIMS = {"fn": (2, 3), "hes1": (3, 7), "hes1log": (3, 7), "hes1log_fixg": (3, 6), "hes1log_fixf": (3, 6), "hiv": (4, 9),
         "ptrans": (5, 6), "lv": (2, 4)}


@dataclass
class OdeSystem:
    """Mirror of ``OdeSystem`` (src/ode_models.jl:5-13): the three callables become ``model_id``."""
    name: str
    model_id: int
    n_dims: int
    thetaSize: int
    thetaLowerBound: np.ndarray = None
    thetaUpperBound: np.ndarray = None

    def __post_init__(self):
        if self.thetaLowerBound is None:
            self.thetaLowerBound = np.full(self.thetaSize, -np.inf)
        if self.thetaUpperBound is None:
            self.thetaUpperBound = np.full(self.thetaSize, np.inf)


def get_ode_system(name: str, n_dims: int | None = None, lower=None, upper=None) -> OdeSystem:
    if name == "lorenz96":
        D, k = int(n_dims or 64), 1
    else:
        D, k = _DIMS[name]
    lo = None if lower is None else np.asarray(lower, dtype=np.float64)
    up = None if upper is None else np.asarray(upper, dtype=np.float64)
    return OdeSystem(name, MODEL_IDS[name], D, k, lo, up)


def fn_system() -> OdeSystem:
    """FitzHugh-Nagumo with the bounds run_scripts/fn_example.jl uses (θ ≥ 0)."""
    return get_ode_system("fn", lower=[0.0, 0.0, 0.0], upper=[np.inf, np.inf, np.inf])


def hes1_system() -> OdeSystem:
    return get_ode_system("hes1", lower=[0.0] * 7, upper=[np.inf] * 7)


def lv_system() -> OdeSystem:
    return get_ode_system("lv", lower=[0.0] * 4, upper=[np.inf] * 4)
