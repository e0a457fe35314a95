"""Kernel constructors (reference: src/kernels.jl:42-50, 74-81).  The reference composes KernelFunctions objects
(variance * Base ∘ ScaleTransform(1/ℓ)); here a kernel is the (family, σ², ℓ) triple the device code consumes."""
from __future__ import annotations

from dataclasses import dataclass

from . import _lib


@dataclass(frozen=True)
class Kernel:
    kind: str            # "matern52" | "rbf" | "matern_nu12" | "matern_nu32" | "matern_nu52" (general Matérn: no time derivatives)
    variance: float
    lengthscale: float

    @property
    def kernel_id(self) -> int:
        return {"matern52": _lib.KERNEL_MATERN52, "rbf": _lib.KERNEL_RBF, "matern_nu12": _lib.KERNEL_MATERN_NU12,
                "matern_nu32": _lib.KERNEL_MATERN_NU32, "matern_nu52": _lib.KERNEL_MATERN_NU52}[self.kind]


def create_rbf_kernel(variance: float, lengthscale: float) -> Kernel:
    """k(t,t') = σ² exp(−(t−t')²/(2ℓ²))   (src/kernels.jl:42-50)"""
    assert variance > 0, "Variance (σ²) must be positive"
    assert lengthscale > 0, "Lengthscale (ℓ) must be positive"
    return Kernel("rbf", float(variance), float(lengthscale))


def create_matern52_kernel(variance: float, lengthscale: float) -> Kernel:
    """k(r) = σ² (1 + √5 r/ℓ + 5r²/(3ℓ²)) exp(−√5 r/ℓ)   (src/kernels.jl:74-81)"""
    assert variance > 0, "Variance (σ²) must be positive"
    assert lengthscale > 0, "Lengthscale (ℓ) must be positive"
    return Kernel("matern52", float(variance), float(lengthscale))


def create_general_matern_kernel(variance: float, lengthscale: float, nu: float) -> Kernel:
    """``variance * MaternKernel(ν=ν) ∘ ScaleTransform(1/ℓ)`` (src/kernels.jl:109-118) for the half-integer orders with a
    closed form (ν = 1/2, 3/2, 5/2).  As in the reference, the GP setup has no analytic time derivatives for this kernel type
    (src/gaussian_process.jl:278-280): ``calculate_gp_covariances`` builds C and takes the zero-derivative fallback."""
    assert variance > 0, "Variance (σ²) must be positive"
    assert lengthscale > 0, "Lengthscale (ℓ) must be positive"
    assert nu > 0, "Smoothness parameter (ν) must be positive"
    kinds = {0.5: "matern_nu12", 1.5: "matern_nu32", 2.5: "matern_nu52"}
    if float(nu) not in kinds:
        raise ValueError("only ν = 1/2, 3/2, 5/2 have a closed form on the device (got %r)" % (nu,))
    return Kernel(kinds[float(nu)], float(variance), float(lengthscale))
