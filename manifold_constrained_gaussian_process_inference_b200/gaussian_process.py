"""GPCov container and calculate_gp_covariances (reference: src/gaussian_process.jl:14-54, 70-74, 219-363).
The matrices are built on the GPU by libmagi_b200 (covariance build, blocked Cholesky, triangular inverse, GEMMs,
band extraction: csrc/setup_kernels.cu); this module only moves them into the reference's field names."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .kernels import Kernel


@dataclass
class GPCov:
    """Field-for-field mirror of ``GPCov`` (src/gaussian_process.jl:14-38).  Dense fields are n×n arrays; the
    three *Band fields are (2b+1)×n diagonal-major tables, T[b + (j−i), i] = A[i, j] (zero outside the matrix)."""
    phi: np.ndarray = field(default_factory=lambda: np.zeros(0))
    tvec: np.ndarray = field(default_factory=lambda: np.zeros(0))
    kernel: Kernel = None
    C: np.ndarray = None
    Cinv: np.ndarray = None
    Cprime: np.ndarray = None
    Cdoubleprime: np.ndarray = None
    mphi: np.ndarray = None
    Kphi: np.ndarray = None
    Kinv: np.ndarray = None
    bandsize: int = 0
    CinvBand: np.ndarray = None
    mphiBand: np.ndarray = None
    KinvBand: np.ndarray = None
    mu: np.ndarray = field(default_factory=lambda: np.zeros(0))
    dotmu: np.ndarray = field(default_factory=lambda: np.zeros(0))
    repaired_pivots: tuple = (0, 0)
    setup_mode: str = "reference_order"

    def band_dense(self, which: str) -> np.ndarray:
        """Dense n×n copy of a band table (zeros outside the band), like ``Matrix(gp_cov.CinvBand)``."""
        T = getattr(self, which)
        b, n = self.bandsize, T.shape[1]
        M = np.zeros((n, n))
        for off in range(-b, b + 1):
            i0, i1 = max(0, -off), min(n, n - off)
            if i1 > i0:
                ii = np.arange(i0, i1)
                M[ii, ii + off] = T[b + off, ii]
        return M


def mat2band(mat_input, l: int, u: int) -> np.ndarray:
    """``mat2band`` (src/gaussian_process.jl:70-74): keeps −u ≤ i−j ≤ l, zero elsewhere (returned dense)."""
    M = np.asarray(mat_input, dtype=np.float64)
    i = np.arange(M.shape[0])[:, None]
    j = np.arange(M.shape[1])[None, :]
    return np.where((i - j <= l) & (j - i <= u), M, 0.0)


_SETUP_MODES = {"reference_order": _lib.SETUP_REFERENCE_ORDER, "stable": _lib.SETUP_STABLE}


def calculate_gp_covariances(gp_cov: GPCov, kernel: Kernel, phi, tvec, bandsize: int, complexity: int = 0,
                             jitter: float = 1e-7, setup_mode: str = "reference_order", device: int = 0) -> None:
    """``calculate_gp_covariances!`` (src/gaussian_process.jl:219-363), computed on the GPU.

    ``complexity < 2`` or a kernel other than Matérn-5/2 / RBF takes the reference's zero-derivative fallback
    (C' = C'' = m = 0, K = εI, Kinv = I/ε; :278-280, :319-331)."""
    L = _lib.lib()
    if not hasattr(L, "magi_gp_covariances"):
        raise _lib.MagiError(-1, "libmagi_b200.so was built without the device GP setup")
    t = np.ascontiguousarray(tvec, dtype=np.float64)
    n = t.shape[0]
    b = int(bandsize)
    ph = np.ascontiguousarray([kernel.variance, kernel.lengthscale], dtype=np.float64)
    dense = [np.zeros((n, n), order="F") for _ in range(7)]
    bands = [np.zeros((2 * b + 1, n)) for _ in range(3)]
    rep = (ctypes.c_int * 2)()
    rc = L.magi_gp_covariances(kernel.kernel_id, _lib.as_dp(ph), _lib.as_dp(t), n, b, float(jitter), int(complexity),
                               _SETUP_MODES[setup_mode], int(device), *[_lib.as_dp(a) for a in dense],
                               *[_lib.as_dp(a) for a in bands], rep)
    _lib.check(rc)
    gp_cov.phi = np.asarray(phi, dtype=np.float64)
    gp_cov.tvec = t
    gp_cov.kernel = kernel
    gp_cov.bandsize = b
    (gp_cov.C, gp_cov.Cinv, gp_cov.Cprime, gp_cov.Cdoubleprime, gp_cov.mphi, gp_cov.Kphi, gp_cov.Kinv) = [np.array(a) for a in dense]
    gp_cov.CinvBand, gp_cov.mphiBand, gp_cov.KinvBand = bands
    gp_cov.mu = np.zeros(n)
    gp_cov.dotmu = np.zeros(n)
    gp_cov.repaired_pivots = (int(rep[0]), int(rep[1]))
    gp_cov.setup_mode = setup_mode
