"""CPU oracle for the MAGI hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference's (MagiJl.jl, pure Julia) per-leapfrog
log-posterior + gradient and of the GP setup that feeds it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` leg may import this module; the product package
(``manifold_constrained_gaussian_process_inference_b200``) never does.

Every function cites the reference file:line it follows (paths relative to
/root/reference).  The reference itself cannot be executed here (Julia is not
installed, no network), so the oracle is pinned the only way available:
against every known answer the reference's own tests hold for this path
(tests/test_oracle_pins.py re-encodes them):

* test/test_likelihoods.jl:62-103,165-179  analytic x/theta-gradient == 5-point FD (rtol 1e-3, atol 1e-4)
* test/test_likelihoods.jl:106-155         missing observation: ll decreases, gradient element moves by exactly +1.0
* test/test_gp.jl                          every GP identity (diag C, C' antisymmetric, C'' diag, m == C'Cinv, K, K*Kinv == I, bands)
* test/test_gp_utils.jl                    band rule
* test/test_kernels.jl:36,73               kernel closed forms
* test/test_ode_models.jl:61,90,120,170,225,244,260,291,326  ODE closed forms

PARITY UNPINNED (no reference test asserts it; third-party arithmetic not under
/root/reference): the numeric value of ll; the sigma-part of the gradient (pinned
only by the formula, likelihoods.jl:229-246); element-wise Cinv/Kinv below
cond*eps; KernelFunctions' pairwise-distance rounding below 1e-8;
PositiveFactorizations' repair of non-positive pivots (``positive_cholesky``
below restates the *recalled* rule and is labelled unverified).

All functions are dtype-generic: pass ``np.longdouble`` arrays to get an
80-bit evaluation that bounds the float64 rounding envelope.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# --------------------------------------------------------------------------
# Kernels  (src/kernels.jl:42-50, 74-81; arithmetic = KernelFunctions 0.10.65,
# not vendored: restated from its published closed forms)
# --------------------------------------------------------------------------

MATERN52 = 0
RBF = 1
MATERN_NU12, MATERN_NU32, MATERN_NU52 = 2, 3, 4      # MaternKernel(ν) of src/kernels.jl:109-118 (no time derivatives in the reference)
KERNEL_IDS = {"matern52": MATERN52, "rbf": RBF, "matern_nu12": MATERN_NU12, "matern_nu32": MATERN_NU32, "matern_nu52": MATERN_NU52}


def kernel_matrix(kernel: int, tvec, variance, lengthscale):
    """``kernelmatrix(variance * Base() ∘ ScaleTransform(1/ℓ), tvec)``
    (call site src/gaussian_process.jl:249).  Inputs are scaled by s = 1/ℓ first
    (ScaleTransform), the distance is taken on the scaled inputs, then
    Matern52: (1 + √5 d + 5 d²/3) exp(−√5 d);  SqExponential: exp(−d²/2)."""
    t = np.asarray(tvec)
    dt = t.dtype.type
    s = dt(1.0) / dt(lengthscale)
    ts = t * s
    diff = ts[:, None] - ts[None, :]
    if kernel in (MATERN52, MATERN_NU52):
        d = np.abs(diff)
        sqrt5 = np.sqrt(dt(5.0))
        base = (dt(1.0) + sqrt5 * d + dt(5.0) * d * d / dt(3.0)) * np.exp(-sqrt5 * d)
    elif kernel == MATERN_NU32:
        d = np.abs(diff)
        sqrt3 = np.sqrt(dt(3.0))
        base = (dt(1.0) + sqrt3 * d) * np.exp(-sqrt3 * d)
    elif kernel == MATERN_NU12:
        base = np.exp(-np.abs(diff))
    elif kernel == RBF:
        base = np.exp(-(diff * diff) / dt(2.0))
    else:
        raise ValueError("unknown kernel id")
    return dt(variance) * base


def kernel_scalar(kernel: int, t1, t2, variance, lengthscale):
    """Scalar k(t, t') -- pins test/test_kernels.jl:36,73."""
    return kernel_matrix(kernel, np.array([t1, t2], dtype=np.float64), variance, lengthscale)[0, 1]


def matern52_derivatives(tvec, variance, lengthscale):
    """C' = ∂k/∂t and C'' = ∂²k/∂t∂t' for Matérn-5/2, in the reference's operation
    order (src/gaussian_process.jl:78-123)."""
    t = np.asarray(tvec)
    dt = t.dtype.type
    n = t.shape[0]
    sqrt5 = np.sqrt(dt(5.0))
    l = dt(lengthscale)
    v = dt(variance)
    l_sq = l * l
    l_cub = l * l * l
    term_div_3l_sq = dt(1.0) / (dt(3.0) * l_sq)
    term_div_3l_cub = dt(1.0) / (dt(3.0) * l_cub)
    t_diff = t[:, None] - t[None, :]
    dist = np.abs(t_diff)
    dist_sq = dist * dist
    sgn = np.sign(t_diff)
    exp_term = np.exp(-sqrt5 * dist / l)
    common = exp_term * (dt(5.0) * dist * term_div_3l_sq + dt(5.0) * sqrt5 * dist_sq * term_div_3l_cub)
    Cp = -sgn * v * common
    term1 = (-sqrt5 / l * exp_term) * (dt(5.0) * dist * term_div_3l_sq + dt(5.0) * sqrt5 * dist_sq * term_div_3l_cub)
    term2 = exp_term * (dt(5.0) * term_div_3l_sq + dt(10.0) * sqrt5 * dist * term_div_3l_cub)
    Cpp = v * (term1 + term2)
    idx = np.arange(n)
    Cp[idx, idx] = dt(0.0)                         # :103
    Cpp[idx, idx] = dt(5.0) * v / (dt(3.0) * l_sq)  # :105
    return Cp, Cpp


def rbf_derivatives(C, tvec, lengthscale):
    """src/gaussian_process.jl:128-154."""
    t = np.asarray(tvec)
    dt = t.dtype.type
    l = dt(lengthscale)
    l_sq = l * l
    l_quad = l_sq * l_sq          # lengthscale^4
    t_diff = t[:, None] - t[None, :]
    t_diff_sq = t_diff * t_diff
    Cp = -C * t_diff / l_sq
    Cpp = C * (dt(1.0) / l_sq - t_diff_sq / l_quad)
    return Cp, Cpp


# --------------------------------------------------------------------------
# Band rule  (src/gaussian_process.jl:70-74; BandedMatrices 1.9.4)
# --------------------------------------------------------------------------

def mat2band(M, l: int, u: int):
    """Dense copy of ``BandedMatrix(M, (l, u))``: keeps entries with −u ≤ i−j ≤ l
    (l sub-diagonals, u super-diagonals), zero elsewhere.
    Pinned by test/test_gp_utils.jl:73-87,117-120,181-184,226-229."""
    M = np.asarray(M)
    n, m = M.shape
    i = np.arange(n)[:, None]
    j = np.arange(m)[None, :]
    keep = (i - j <= l) & (j - i <= u)
    return np.where(keep, M, M.dtype.type(0.0))


def band_storage(M, b: int):
    """(2b+1) × n diagonal-major table: T[b + (j − i), i] = M[i, j] for |i−j| ≤ b
    (zero where j is out of range).  This is the layout ``magi_set_band_tables`` /
    ``magi_get_matrix`` exchange (row-offset major, time fastest)."""
    M = np.asarray(M)
    n = M.shape[0]
    T = np.zeros((2 * b + 1, n), dtype=M.dtype)
    for off in range(-b, b + 1):
        i0, i1 = max(0, -off), min(n, n - off)
        if i1 > i0:
            ii = np.arange(i0, i1)
            T[b + off, ii] = M[ii, ii + off]
    return T


def band_from_storage(T, b: int):
    """Inverse of ``band_storage`` (dense n×n with zeros outside the band)."""
    T = np.asarray(T)
    n = T.shape[1]
    M = np.zeros((n, n), dtype=T.dtype)
    for off in range(-b, b + 1):
        i0, i1 = max(0, -off), min(n, n - off)
        if i1 > i0:
            ii = np.arange(i0, i1)
            M[ii, ii + off] = T[b + off, ii]
    return M


def band_matvec(T, b: int, x, transpose: bool = False):
    """y = A x (or Aᵀ x) for A given as a diagonal-major band table.  Per-row
    accumulation runs over increasing column index, the order ``dgbmv`` produces
    (call sites src/likelihoods.jl:129,132,133,192)."""
    T = np.asarray(T)
    n = T.shape[1]
    y = np.zeros(n, dtype=np.result_type(T.dtype, np.asarray(x).dtype))
    if not transpose:
        for off in range(-b, b + 1):
            i0, i1 = max(0, -off), min(n, n - off)
            if i1 > i0:
                y[i0:i1] += T[b + off, i0:i1] * x[i0 + off:i1 + off]
    else:
        # (Aᵀ x)[j] = Σ_i A[i, j] x[i];  A[i, j] = T[b + j − i, i]; increasing i
        for off in range(b, -b - 1, -1):          # i = j − off, increasing i ⇔ decreasing off
            i0, i1 = max(0, -off), min(n, n - off)
            if i1 > i0:
                y[i0 + off:i1 + off] += T[b + off, i0:i1] * x[i0:i1]
    return y


# --------------------------------------------------------------------------
# Factorisations (PositiveFactorizations 0.2.4 + LAPACK potri; not vendored)
# --------------------------------------------------------------------------

def positive_cholesky(A, repair: bool = True):
    """Lower factor L of the *upper-triangle-authoritative* symmetric matrix A
    (``Symmetric(·)`` reads the upper triangle, src/gaussian_process.jl:257,306).

    For a numerically SPD input this is a plain unpivoted Cholesky.  For a
    non-positive pivot the reference's ``cholesky(Positive, ·)`` does not throw;
    the rule restated here is RECALLED from PositiveFactorizations (source not
    available; UNVERIFIED, parity unpinned): pivot p_j ≤ tol·... is replaced by
    |p_j| (sign flipped) and a near-zero pivot by the tolerance, then the
    elimination continues as if D = +I.  Returns (L, n_repaired)."""
    A = np.asarray(A)
    dt = A.dtype.type
    n = A.shape[0]
    S = np.triu(A) + np.triu(A, 1).T
    L = np.zeros_like(S)
    W = S.copy()
    repaired = 0
    eps = np.finfo(A.dtype).eps
    tol = dt(n) * eps * (np.max(np.abs(np.diag(S))) if n else dt(0.0))
    for j in range(n):
        p = W[j, j]
        if not (p > tol):
            if not repair:
                raise np.linalg.LinAlgError("matrix is not positive definite (pivot %d = %r)" % (j, p))
            repaired += 1
            p = abs(p) if abs(p) > tol else (tol if tol > 0 else dt(1.0))
        r = np.sqrt(p)
        L[j, j] = r
        if j + 1 < n:
            c = W[j + 1:, j] / r
            L[j + 1:, j] = c
            W[j + 1:, j + 1:] -= np.outer(c, c)
    return L, repaired


def inverse_from_cholesky(L):
    """``inv(cholesky_object)``: potri on the factor then mirror, so the result is
    exactly symmetric (src/gaussian_process.jl:296,318)."""
    L = np.asarray(L)
    n = L.shape[0]
    if L.dtype == np.float64:
        from scipy.linalg import lapack
        inv, info = lapack.dpotri(L, lower=1)
        if info != 0:
            raise np.linalg.LinAlgError("dpotri info=%d" % info)
        inv = np.tril(inv) + np.tril(inv, -1).T
        return inv
    # extended precision: explicit triangular inverse, then LinvᵀLinv
    Linv = np.zeros_like(L)
    I = np.eye(n, dtype=L.dtype)
    for j in range(n):
        col = I[:, j].copy()
        for i in range(j, n):
            col[i] = (col[i] - np.dot(L[i, j:i], col[j:i])) / L[i, i]
        Linv[:, j] = col
    inv = Linv.T @ Linv
    return np.tril(inv) + np.tril(inv, -1).T


# --------------------------------------------------------------------------
# GPCov + calculate_gp_covariances!  (src/gaussian_process.jl:14-54, 219-363)
# --------------------------------------------------------------------------

@dataclass
class GPCov:
    phi: np.ndarray = None
    tvec: np.ndarray = None
    kernel: int = MATERN52
    C: np.ndarray = None
    Cinv: np.ndarray = None
    Cprime: np.ndarray = None
    Cdoubleprime: np.ndarray = None
    mphi: np.ndarray = None
    Kphi: np.ndarray = None        # the *jittered* K (:306-307)
    Kinv: np.ndarray = None
    bandsize: int = 0
    CinvBand: np.ndarray = None    # diagonal-major (2b+1)×n tables (band_storage)
    mphiBand: np.ndarray = None
    KinvBand: np.ndarray = None
    repaired_pivots: tuple = (0, 0)
    setup_mode: str = "reference_order"


def calculate_gp_covariances(kernel: int, phi, tvec, bandsize: int, complexity: int = 2,
                             jitter: float = 1e-7, setup_mode: str = "reference_order",
                             dtype=np.float64) -> GPCov:
    """src/gaussian_process.jl:219-363.

    ``setup_mode="reference_order"`` follows the reference: Cinv = inv(chol(C+εI)),
    m = C'·Cinv, K = Sym_upper(C'' − m·C'ᵀ + εI), Kinv = inv(chol(K)).
    ``setup_mode="stable"`` is the triangular-solve route of SURVEY.md F11
    (W = L⁻¹C'ᵀ, K = C'' − WᵀW + εI, m = (L⁻ᵀW)ᵀ): same mathematics, PD by
    construction; NOT what the reference computes in floating point.

    ``kernel`` other than Matérn-5/2 / RBF or ``complexity < 2`` takes the
    reference's zero-derivative fallback (:278-280, :319-331)."""
    t = np.asarray(tvec, dtype=dtype)
    dt = t.dtype.type
    n = t.shape[0]
    variance, lengthscale = dt(phi[0]), dt(phi[1])
    g = GPCov(phi=np.asarray(phi, dtype=dtype), tvec=t, kernel=kernel, bandsize=bandsize, setup_mode=setup_mode)
    eps = dt(jitter)
    I = np.eye(n, dtype=dtype)
    g.C = kernel_matrix(kernel, t, variance, lengthscale)
    Cj = g.C + eps * I                                           # :257
    derivatives = False
    g.Cprime = np.zeros((n, n), dtype=dtype)
    g.Cdoubleprime = np.zeros((n, n), dtype=dtype)
    if complexity >= 2:
        if kernel == MATERN52:
            g.Cprime, g.Cdoubleprime = matern52_derivatives(t, variance, lengthscale)   # :273
            derivatives = True
        elif kernel == RBF:
            g.Cprime, g.Cdoubleprime = rbf_derivatives(g.C, t, lengthscale)            # :276
            derivatives = True
    L, rep_c = positive_cholesky(Cj)                              # :295
    rep_k = 0
    if derivatives and np.any(g.Cprime != 0) and np.any(g.Cdoubleprime != 0):          # :299
        if setup_mode == "reference_order":
            g.Cinv = inverse_from_cholesky(L)                     # :296
            g.mphi = g.Cprime @ g.Cinv                            # :302
            Kd = g.Cdoubleprime - g.mphi @ g.Cprime.T             # :304
            Kj = Kd + eps * I
            Kj = np.triu(Kj) + np.triu(Kj, 1).T                   # :306 Symmetric(upper)
        elif setup_mode == "stable":
            g.Cinv = inverse_from_cholesky(L)
            W = _solve_lower(L, g.Cprime.T)                       # W = L⁻¹ C'ᵀ
            Kj = g.Cdoubleprime - W.T @ W + eps * I
            Kj = np.triu(Kj) + np.triu(Kj, 1).T
            g.mphi = _solve_upper(L.T, W).T                       # m = (L⁻ᵀ W)ᵀ = C' (C+εI)⁻¹
        else:
            raise ValueError("setup_mode")
        g.Kphi = Kj                                               # :307
        Lk, rep_k = positive_cholesky(Kj)                         # :317
        g.Kinv = inverse_from_cholesky(Lk)                        # :318
    else:                                                         # :319-331
        g.Cinv = inverse_from_cholesky(L)
        g.mphi = np.zeros((n, n), dtype=dtype)
        g.Kphi = eps * I
        Lk, rep_k = positive_cholesky(g.Kphi)
        g.Kinv = inverse_from_cholesky(Lk)
    g.repaired_pivots = (rep_c, rep_k)
    b = bandsize
    g.CinvBand = band_storage(g.Cinv, b)                          # :358-360
    g.mphiBand = band_storage(g.mphi, b)
    g.KinvBand = band_storage(g.Kinv, b)
    return g


def _solve_lower(L, B):
    if L.dtype == np.float64:
        from scipy.linalg import solve_triangular
        return solve_triangular(L, B, lower=True)
    n = L.shape[0]
    X = np.array(B, dtype=L.dtype, copy=True)
    for i in range(n):
        X[i] = (X[i] - L[i, :i] @ X[:i]) / L[i, i]
    return X


def _solve_upper(U, B):
    if U.dtype == np.float64:
        from scipy.linalg import solve_triangular
        return solve_triangular(U, B, lower=False)
    n = U.shape[0]
    X = np.array(B, dtype=U.dtype, copy=True)
    for i in range(n - 1, -1, -1):
        X[i] = (X[i] - U[i, i + 1:] @ X[i + 1:]) / U[i, i]
    return X


# --------------------------------------------------------------------------
# ODE models (src/ode_models.jl).  f(X, θ) → (n, D);  dfdx → (n, D, D) with
# [i, p, j] = ∂f_p/∂x_j;  dfdtheta → (n, D, k) with [i, p, q] = ∂f_p/∂θ_q.
# Model ids are the C-ABI's ``ode_model_id``.
# --------------------------------------------------------------------------

@dataclass
class OdeModel:
    model_id: int
    name: str
    n_dims: int
    n_params: int
    f: callable
    dfdx: callable = None
    dfdtheta: callable = None
    in_reference: bool = True
    consts: tuple = ()


def _fn_f(X, th):
    V, R = X[:, 0], X[:, 1]
    a, b, c = th[0], th[1], th[2]
    one, three = X.dtype.type(1.0), X.dtype.type(3.0)
    return np.stack([c * (V - (V * V * V) / three + R),        # src/ode_models.jl:44
                     -one / c * (V - a + b * R)], axis=1)      # :45


def _fn_dx(X, th):
    V = X[:, 0]
    a, b, c = th[0], th[1], th[2]
    one = X.dtype.type(1.0)
    J = np.zeros((X.shape[0], 2, 2), dtype=X.dtype)
    J[:, 0, 0] = c * (one - V * V)      # :254
    J[:, 0, 1] = c                      # :256
    J[:, 1, 0] = -one / c               # :258
    J[:, 1, 1] = -b / c                 # :260
    return J


def _fn_dth(X, th):
    V, R = X[:, 0], X[:, 1]
    a, b, c = th[0], th[1], th[2]
    one, three = X.dtype.type(1.0), X.dtype.type(3.0)
    Jp = np.zeros((X.shape[0], 2, 3), dtype=X.dtype)
    Jp[:, 0, 2] = V - (V * V * V) / three + R           # :288
    Jp[:, 1, 0] = one / c                               # :292
    Jp[:, 1, 1] = -R / c                                # :294
    Jp[:, 1, 2] = (one / (c * c)) * (V - a + b * R)     # :296
    return Jp


def _hes1_f(X, p):
    P, M, H = X[:, 0], X[:, 1], X[:, 2]
    one = X.dtype.type(1.0)
    return np.stack([-p[0] * P * H + p[1] * M - p[2] * P,                       # :66
                     -p[3] * M + p[4] / (one + P * P),                          # :67
                     -p[0] * P * H + p[5] / (one + P * P) - p[6] * H], axis=1)   # :68


def _hes1_dx(X, p):
    P, H = X[:, 0], X[:, 2]
    one, two = X.dtype.type(1.0), X.dtype.type(2.0)
    opp = one + P * P
    J = np.zeros((X.shape[0], 3, 3), dtype=X.dtype)
    J[:, 0, 0] = -p[0] * H - p[2]
    J[:, 0, 1] = p[1]
    J[:, 0, 2] = -p[0] * P
    J[:, 1, 0] = -p[4] * (two * P) / (opp * opp)                 # :326
    J[:, 1, 1] = -p[3]
    J[:, 2, 0] = -p[0] * H - p[5] * (two * P) / (opp * opp)      # :331
    J[:, 2, 2] = -p[0] * P - p[6]
    return J


def _hes1_dth(X, p):
    P, M, H = X[:, 0], X[:, 1], X[:, 2]
    one = X.dtype.type(1.0)
    opp = one + P * P
    Jp = np.zeros((X.shape[0], 3, 7), dtype=X.dtype)
    Jp[:, 0, 0] = -P * H
    Jp[:, 0, 1] = M
    Jp[:, 0, 2] = -P
    Jp[:, 1, 3] = -M
    Jp[:, 1, 4] = one / opp
    Jp[:, 2, 0] = -P * H
    Jp[:, 2, 5] = one / opp
    Jp[:, 2, 6] = -H
    return Jp


def _hes1log_core(X, p1, p2, p3, p4, p5, p6, p7):
    P, M, H = np.exp(X[:, 0]), np.exp(X[:, 1]), np.exp(X[:, 2])
    one = X.dtype.type(1.0)
    opp = one + P * P
    return np.stack([-p1 * H + p2 * M / P - p3,           # :97
                     -p4 + p5 / (opp * M),                # :99
                     -p1 * P + p6 / (opp * H) - p7], axis=1)   # :101


def _hes1log_f(X, p):
    return _hes1log_core(X, p[0], p[1], p[2], p[3], p[4], p[5], p[6])


def _hes1log_fixg_f(X, p):      # :116-135, γ fixed at 0.3
    return _hes1log_core(X, p[0], p[1], p[2], p[3], p[4], p[5], X.dtype.type(0.3))


def _hes1log_fixf_f(X, p):      # :147-165, f fixed at 20.0
    return _hes1log_core(X, p[0], p[1], p[2], p[3], p[4], X.dtype.type(20.0), p[5])


def _hiv_f(X, p):               # :178-207
    T, Tm, Tw, Tmw = np.exp(X[:, 0]), np.exp(X[:, 1]), np.exp(X[:, 2]), np.exp(X[:, 3])
    dt = X.dtype.type
    sf = dt(1e-6)
    return np.stack([
        p[0] - sf * p[1] * Tm - sf * p[2] * Tw - sf * p[3] * Tmw,
        p[6] + sf * p[1] * T - sf * p[4] * Tw + sf * dt(0.25) * p[3] * Tmw * T / Tm,
        p[7] + sf * p[2] * T - sf * p[5] * Tm + sf * dt(0.25) * p[3] * Tmw * T / Tw,
        p[8] + dt(0.5) * sf * p[3] * T + (sf * p[4] + sf * p[5]) * Tw * Tm / Tmw], axis=1)


def _ptrans_f(X, p):            # :219-233
    S, R, RS, RPP = X[:, 0], X[:, 2], X[:, 3], X[:, 4]
    return np.stack([
        -p[0] * S - p[1] * S * R + p[2] * RS,
        p[0] * S,
        -p[1] * S * R + p[2] * RS + p[4] * RPP / (p[5] + RPP),
        p[1] * S * R - p[2] * RS - p[3] * RS,
        p[3] * RS - p[4] * RPP / (p[5] + RPP)], axis=1)


# -- models NOT in the reference; oracle = derivation, FD-checked in tests --

def _lv_f(X, th):
    x, y = X[:, 0], X[:, 1]
    al, be, de, ga = th[0], th[1], th[2], th[3]
    return np.stack([al * x - be * x * y, de * x * y - ga * y], axis=1)


def _lv_dx(X, th):
    x, y = X[:, 0], X[:, 1]
    al, be, de, ga = th[0], th[1], th[2], th[3]
    J = np.zeros((X.shape[0], 2, 2), dtype=X.dtype)
    J[:, 0, 0] = al - be * y
    J[:, 0, 1] = -be * x
    J[:, 1, 0] = de * y
    J[:, 1, 1] = de * x - ga
    return J


def _lv_dth(X, th):
    x, y = X[:, 0], X[:, 1]
    Jp = np.zeros((X.shape[0], 2, 4), dtype=X.dtype)
    Jp[:, 0, 0] = x
    Jp[:, 0, 1] = -x * y
    Jp[:, 1, 2] = x * y
    Jp[:, 1, 3] = -y
    return Jp


def _l96_f(X, th):
    # ẋ_i = (x_{i+1} − x_{i−2}) x_{i−1} − x_i + F, cyclic
    xp1 = np.roll(X, -1, axis=1)
    xm1 = np.roll(X, 1, axis=1)
    xm2 = np.roll(X, 2, axis=1)
    return (xp1 - xm2) * xm1 - X + th[0]


def _l96_dx(X, th):
    n, D = X.shape
    J = np.zeros((n, D, D), dtype=X.dtype)
    for i in range(D):
        ip1, im1, im2 = (i + 1) % D, (i - 1) % D, (i - 2) % D
        J[:, i, ip1] += X[:, im1]
        J[:, i, im2] += -X[:, im1]
        J[:, i, im1] += X[:, ip1] - X[:, im2]
        J[:, i, i] += -1.0
    return J


def _l96_dth(X, th):
    return np.ones((X.shape[0], X.shape[1], 1), dtype=X.dtype)


MODEL_FN, MODEL_HES1, MODEL_HES1LOG, MODEL_HES1LOG_FIXG, MODEL_HES1LOG_FIXF, MODEL_HIV, MODEL_PTRANS, MODEL_LV, MODEL_L96 = range(9)


def get_model(model_id: int, n_dims: int | None = None) -> OdeModel:
    if model_id == MODEL_FN:
        return OdeModel(MODEL_FN, "fn", 2, 3, _fn_f, _fn_dx, _fn_dth)
    if model_id == MODEL_HES1:
        return OdeModel(MODEL_HES1, "hes1", 3, 7, _hes1_f, _hes1_dx, _hes1_dth)
    if model_id == MODEL_HES1LOG:
        return OdeModel(MODEL_HES1LOG, "hes1log", 3, 7, _hes1log_f)
    if model_id == MODEL_HES1LOG_FIXG:
        return OdeModel(MODEL_HES1LOG_FIXG, "hes1log_fixg", 3, 6, _hes1log_fixg_f)
    if model_id == MODEL_HES1LOG_FIXF:
        return OdeModel(MODEL_HES1LOG_FIXF, "hes1log_fixf", 3, 6, _hes1log_fixf_f)
    if model_id == MODEL_HIV:
        return OdeModel(MODEL_HIV, "hiv", 4, 9, _hiv_f)
    if model_id == MODEL_PTRANS:
        return OdeModel(MODEL_PTRANS, "ptrans", 5, 6, _ptrans_f)
    if model_id == MODEL_LV:
        return OdeModel(MODEL_LV, "lv", 2, 4, _lv_f, _lv_dx, _lv_dth, in_reference=False)
    if model_id == MODEL_L96:
        return OdeModel(MODEL_L96, "lorenz96", int(n_dims or 64), 1, _l96_f, _l96_dx, _l96_dth, in_reference=False)
    raise ValueError("unknown ode_model_id %r" % (model_id,))


# --------------------------------------------------------------------------
# log_likelihood_and_gradient_banded  (src/likelihoods.jl:43-257)
# --------------------------------------------------------------------------

def log_likelihood_and_gradient_banded(xlatent, theta, sigma, yobs, covs, model: OdeModel,
                                       prior_temperature=(1.0, 1.0, 1.0)):
    """Returns (ll, grad[nD + k + D]) ordered [vec(gX) (column-major, time fastest); gθ; gσ].

    ``covs`` is a list of D ``GPCov`` (only bandsize and the three band tables are
    read, as in the reference).  Term order and β divisions follow
    src/likelihoods.jl:139-151 (value) and :168-247 (gradient)."""
    X = np.asarray(xlatent)
    dt = X.dtype.type
    n, D = X.shape
    th = np.asarray(theta, dtype=X.dtype)
    sig = np.asarray(sigma, dtype=X.dtype)
    Y = np.asarray(yobs)
    k = th.shape[0]
    beta = [dt(b) for b in prior_temperature]
    if Y.shape != (n, D):
        raise ValueError("Dimensions of yobs do not match xlatent")       # :63
    if sig.shape[0] != D or len(covs) != D or len(beta) != 3:
        raise ValueError("argument size mismatch")                        # :66-74
    half = dt(0.5)
    F = model.f(X, th)                                                    # :89-95
    sigma_sq = sig * sig                                                  # :102
    ll = dt(0.0)
    Ke_all = np.zeros((n, D), dtype=X.dtype)
    Cx_all = np.zeros((n, D), dtype=X.dtype)
    e0_all = np.zeros((n, D), dtype=X.dtype)
    fin_all = np.zeros((n, D), dtype=bool)
    two_pi = dt(2.0) * dt(np.pi) if X.dtype != np.longdouble else dt(2.0) * np.longdouble(math.pi)
    for d in range(D):                                                    # :111-152
        g = covs[d]
        b = g.bandsize
        xd = X[:, d]
        fin = np.isfinite(Y[:, d].astype(np.float64))                     # :123
        e0 = np.where(fin, xd - np.where(fin, Y[:, d], 0).astype(X.dtype), dt(0.0))   # :122,125
        nobs = int(fin.sum())
        mx = band_matvec(g.mphiBand.astype(X.dtype, copy=False), b, xd)   # :129
        e = F[:, d] - mx                                                  # :130
        Ke = band_matvec(g.KinvBand.astype(X.dtype, copy=False), b, e)    # :132
        Cx = band_matvec(g.CinvBand.astype(X.dtype, copy=False), b, xd)   # :133
        Ke_all[:, d], Cx_all[:, d], e0_all[:, d], fin_all[:, d] = Ke, Cx, e0, fin
        ll_obs = -half * np.dot(e0[fin], e0[fin]) / sigma_sq[d]           # :139
        if nobs > 0:
            ll_obs -= half * dt(nobs) * np.log(two_pi * sigma_sq[d])      # :141
        ll += ll_obs / beta[2]                                            # :143
        ll += (-half * np.dot(e, Ke)) / beta[0]                           # :146-147
        ll += (-half * np.dot(xd, Cx)) / beta[1]                          # :150-151
    gX = np.zeros((n, D), dtype=X.dtype)
    gth = np.zeros(k, dtype=X.dtype)
    gsig = np.zeros(D, dtype=X.dtype)
    if model.dfdx is None or model.dfdtheta is None:
        # the reference ships no Jacobians for this model (src/ode_models.jl:83-233): value only, gradient left NaN
        return ll, np.full(n * D + k + D, np.nan, dtype=X.dtype)
    Jx = model.dfdx(X, th)            # (n, D, D); the reference re-evaluates it D times (:199-209) -- same values
    Jp = model.dfdtheta(X, th)        # (n, D, k)
    for d in range(D):                                                    # :168-247
        g = covs[d]
        b = g.bandsize
        fin = fin_all[:, d]
        gX[fin, d] -= (e0_all[fin, d] / sigma_sq[d]) / beta[2]            # :177-181
        gX[:, d] -= Cx_all[:, d] / beta[1]                                # :185-187
        mt = band_matvec(g.mphiBand.astype(X.dtype, copy=False), b, Ke_all[:, d], transpose=True)   # :192
        gX[:, d] += mt / beta[0]                                          # :193-195
        kfe = Ke_all[:, d] / beta[0]                                      # :201
        for j in range(D):
            gX[:, j] -= Jx[:, d, j] * kfe                                 # :214-216
        for q in range(k):
            # sequential accumulation over time, as the reference's scalar loop (:219-221)
            gth[q] -= _seq_sum(Jp[:, d, q] * kfe)
        if sig[d] > 0:                                                    # :229
            sse = _seq_sum(e0_all[fin, d] * e0_all[fin, d])               # :230-237
            npts = int(fin.sum())
            if npts > 0:
                gsig[d] += (sse / sigma_sq[d] - dt(npts)) / (sig[d] * beta[2])   # :243
    grad = np.concatenate([gX.reshape(-1, order="F"), gth, gsig])
    return ll, grad


def _seq_sum(v):
    """Left-to-right sum (the reference accumulates in scalar loops)."""
    s = v.dtype.type(0.0)
    for a in v:
        s += a
    return s


# --------------------------------------------------------------------------
# MagiTarget + LogDensityProblems interface
# (src/logdensityproblems_interface.jl:33-45, 53-70, 79-101, 111-166, 176-267)
# --------------------------------------------------------------------------

@dataclass
class MagiTarget:
    yobs: np.ndarray
    gp_cov_all_dims: list
    model: OdeModel
    sigma_init: np.ndarray
    prior_temperature: tuple
    n_times: int
    n_dims: int
    n_params_ode: int
    sigma_is_fixed: bool
    dtype: type = np.float64


def dimension(target: MagiTarget) -> int:                                # :53-61
    dim = target.n_times * target.n_dims + target.n_params_ode
    if not target.sigma_is_fixed:
        dim += target.n_dims
    return dim


def _unpack(target, params):                                              # :79-101
    n, D, k = target.n_times, target.n_dims, target.n_params_ode
    X = params[:n * D].reshape((n, D), order="F")
    th = params[n * D:n * D + k]
    ls = None if target.sigma_is_fixed else params[n * D + k:]
    return X, th, ls


def logdensity_and_gradient(target: MagiTarget, params):
    """(:176-267).  Returns (ll, grad[P]).  Guards: wrong length → (−Inf, NaN…);
    invalid fixed σ → (−Inf, NaN…); non-finite likelihood/gradient → (−Inf, 0…);
    non-finite final gradient → (ll, 0…)."""
    P = dimension(target)
    params = np.asarray(params, dtype=target.dtype)
    dt = params.dtype.type
    if params.shape[0] != P:
        return -np.inf, np.full(P, np.nan)                                # :179-182
    X, th, ls = _unpack(target, params)
    D = target.n_dims
    prior = dt(0.0)
    if target.sigma_is_fixed:
        sigma = np.asarray(target.sigma_init, dtype=target.dtype)
        if np.any(~np.isfinite(sigma.astype(np.float64)) | (sigma <= 0)):
            return -np.inf, np.full(P, np.nan)                            # :192-195
        gjac = np.zeros(D, dtype=target.dtype)
    else:
        cl = np.clip(ls, dt(-15.0), dt(15.0))                             # :200
        sigma = np.exp(cl)                                                # :201
        prior = _seq_sum(cl)                                              # :206
        gjac = np.ones(D, dtype=target.dtype)                             # :208
    ll, g = log_likelihood_and_gradient_banded(X, th, sigma, target.yobs, target.gp_cov_all_dims,
                                               target.model, target.prior_temperature)     # :215-219
    if not np.isfinite(np.float64(ll)) or not np.all(np.isfinite(g.astype(np.float64))):
        return -np.inf, np.zeros(P)                                       # :222-226
    n_xt = target.n_times * D + target.n_params_ode
    final = np.zeros(P, dtype=target.dtype)
    final[:n_xt] = g[:n_xt]                                               # :233
    total = ll
    if not target.sigma_is_fixed:
        total = total + prior                                             # :238
        final[n_xt:] = g[n_xt:] * sigma + gjac                            # :249-253
    if not np.all(np.isfinite(final.astype(np.float64))):
        return total, np.zeros(P)                                         # :260-264
    return total, final


def logdensity(target: MagiTarget, params):
    """(:111-166) value-only variant: same computation, gradient discarded; any
    non-finite total → −Inf."""
    P = dimension(target)
    params = np.asarray(params, dtype=target.dtype)
    if params.shape[0] != P:
        return -np.inf
    X, th, ls = _unpack(target, params)
    dt = params.dtype.type
    prior = dt(0.0)
    if target.sigma_is_fixed:
        sigma = np.asarray(target.sigma_init, dtype=target.dtype)
        if np.any(~np.isfinite(sigma.astype(np.float64)) | (sigma <= 0)):
            return -np.inf
    else:
        cl = np.clip(ls, dt(-15.0), dt(15.0))
        sigma = np.exp(cl)
        prior = _seq_sum(cl)
    ll, _ = log_likelihood_and_gradient_banded(X, th, sigma, target.yobs, target.gp_cov_all_dims,
                                               target.model, target.prior_temperature)
    total = ll + prior if not target.sigma_is_fixed else ll
    if not np.isfinite(np.float64(total)):
        return -np.inf
    return total


def make_target(yobs, covs, model_id, sigma_init, prior_temperature, sigma_is_fixed, n_dims=None, dtype=np.float64):
    yobs = np.asarray(yobs)
    n, D = yobs.shape
    model = get_model(model_id, D)
    return MagiTarget(yobs=yobs, gp_cov_all_dims=covs, model=model, sigma_init=np.asarray(sigma_init, dtype=dtype),
                      prior_temperature=tuple(prior_temperature), n_times=n, n_dims=D,
                      n_params_ode=model.n_params, sigma_is_fixed=bool(sigma_is_fixed), dtype=dtype)


# --------------------------------------------------------------------------
# GP hyper-parameter initialisation objective (src/initialization.jl:72-176)
# --------------------------------------------------------------------------

def negative_log_marginal_likelihood(log_params, y_obs_dim, t_obs, kernel: int, jitter: float = 1e-6):
    """0.5 (log|K_phi + (sigma^2 + jitter) I| + y^T (.)^-1 y + N log 2 pi) on the non-NaN observations; Inf for invalid
    parameters, no data or a non-finite result."""
    variance, lengthscale, sigma = (float(np.exp(v)) for v in log_params)
    if not all(np.isfinite([variance, lengthscale, sigma])) or min(variance, lengthscale, sigma) <= 0:
        return np.inf
    y = np.asarray(y_obs_dim, dtype=np.float64)
    ok = ~np.isnan(y)
    if not ok.any():
        return np.inf
    ys, ts = y[ok], np.asarray(t_obs, dtype=np.float64)[ok]
    n = ys.shape[0]
    K = kernel_matrix(kernel, ts, variance, lengthscale) + (sigma * sigma + jitter) * np.eye(n)     # :128
    L, _ = positive_cholesky(K)                                                                  # :135
    logdet = 2.0 * np.sum(np.log(np.diag(L)))                                                    # :138
    z = _solve_lower(L, ys)
    v = 0.5 * (logdet + float(z @ z) + n * np.log(2.0 * np.pi))                                  # :150
    return v if np.isfinite(v) else np.inf
