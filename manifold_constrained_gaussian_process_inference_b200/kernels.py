"""Kernel constructors (reference: src/kernels.jl:42-50, 74-81).  The reference composes KernelFunctions objects
(variance * Base ∘ ScaleTransform(1/ℓ)); here a kernel is the (family, σ², ℓ) triple the device code consumes."""
from __future__ import annotations

from dataclasses import dataclass

from . import _lib


@dataclass(frozen=True)
class Kernel:
    kind: str            # "matern52" | "rbf"
    variance: float
    lengthscale: float

    @property
    def kernel_id(self) -> int:
        return {"matern52": _lib.KERNEL_MATERN52, "rbf": _lib.KERNEL_RBF}[self.kind]


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
