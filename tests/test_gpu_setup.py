"""Device GP setup (K3-K6) against the reference's own test battery (test/test_gp.jl, test/test_gp_utils.jl) and,
element-wise, against the oracle in the benign regime (element-wise parity of Cinv/Kinv is only
defined where cond*eps is small; at BASELINE scale the checks are backward-error identities)."""
import numpy as np
import pytest

from oracle import magi_oracle as mo
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _setup(pkg, kernel, var, ell, tvec, b, jitter, complexity=2, mode="reference_order"):
    g = pkg.GPCov()
    k = pkg.create_matern52_kernel(var, ell) if kernel == "matern52" else pkg.create_rbf_kernel(var, ell)
    pkg.calculate_gp_covariances(g, k, [var, ell], tvec, b, complexity=complexity, jitter=jitter, setup_mode=mode)
    return g


@pytest.mark.parametrize("kernel,var,ell,tvec,b,eps", [
    ("matern52", 1.5, 0.8, np.arange(0.0, 1.0 + 1e-9, 0.2), 2, 1e-6),     # test/test_gp.jl:40-250
    ("rbf", 2.0, 1.2, np.arange(0.0, 1.0 + 1e-9, 0.2), 2, 1e-6),          # test/test_gp.jl:255-362
    ("rbf", 2.5, 0.3, np.linspace(0.0, 1.0, 5), 4, 1e-6),                 # b = N-1
    ("matern52", 2.0, 1.5, np.linspace(0.0, 20.0, 41), 20, 1e-6),         # benign n=41
    ("matern52", 2.0, 0.5, np.linspace(0.0, 20.0, 201), 20, 1e-6),        # benign n=201, short lengthscale
])
def test_gp_identities_and_oracle(pkg, kernel, var, ell, tvec, b, eps):
    n = len(tvec)
    g = _setup(pkg, kernel, var, ell, tvec, b, eps)
    I = np.eye(n)
    assert np.allclose(np.diag(g.C), var, atol=1e-9)                                        # :75
    assert np.max(np.abs((g.C + eps * I) @ g.Cinv - I)) < 1e-6                               # :83
    assert np.allclose(g.Cprime, -g.Cprime.T, atol=1e-12) and np.all(np.diag(g.Cprime) == 0)  # :101-111
    assert np.allclose(g.Cdoubleprime, g.Cdoubleprime.T, atol=1e-12)                         # :142
    expect_diag = 5 * var / (3 * ell ** 2) if kernel == "matern52" else var / ell ** 2
    assert np.allclose(np.diag(g.Cdoubleprime), expect_diag, rtol=1e-12)                     # :147 / :330
    assert np.allclose(g.mphi, g.Cprime @ g.Cinv, atol=1e-7 * max(1.0, np.abs(g.mphi).max()))  # :162
    Kexp = g.Cdoubleprime - g.mphi @ g.Cprime.T + eps * I
    Kexp = np.triu(Kexp) + np.triu(Kexp, 1).T
    assert np.allclose(g.Kphi, Kexp, atol=1e-9 * max(1.0, np.abs(Kexp).max()))               # :171
    assert np.array_equal(g.Kphi, g.Kphi.T) and np.array_equal(g.Cinv, g.Cinv.T) and np.array_equal(g.Kinv, g.Kinv.T)
    assert np.max(np.abs(g.Kphi @ g.Kinv - I)) < 1e-6 * max(1.0, np.linalg.cond(g.Kphi) * 1e-10)   # :197,204
    for dense, band in ((g.Cinv, "CinvBand"), (g.mphi, "mphiBand"), (g.Kinv, "KinvBand")):   # :248-250, test_gp_utils.jl
        assert np.array_equal(g.band_dense(band), pkg.mat2band(dense, b, b))
    # element-wise against the oracle (same operation order; benign regime)
    o = mo.calculate_gp_covariances(mo.KERNEL_IDS[kernel], [var, ell], tvec, b, complexity=2, jitter=eps)
    assert g.repaired_pivots == (0, 0) and o.repaired_pivots == (0, 0)
    for name, tol in (("C", 1e-14), ("Cprime", 1e-13), ("Cdoubleprime", 1e-13)):
        a, r = getattr(g, name), getattr(o, name)
        assert np.max(np.abs(a - r)) <= tol * max(1.0, np.abs(r).max()), name
    # error model: Cinv and m carry cond(C)*eps; K = C'' - m C'^T is a cancellation, so its ABSOLUTE error is
    # cond(C)*eps*|C''| whatever the size of K; Kinv inherits cond(K) times the relative error of K.
    condC = np.linalg.cond(o.C + eps * I)
    condK = np.linalg.cond(o.Kphi)
    u = 50 * np.finfo(float).eps
    for name in ("Cinv", "mphi"):
        a, r = getattr(g, name), getattr(o, name)
        rel = np.max(np.abs(a - r)) / np.abs(r).max()
        assert rel <= u * condC, "%s rel err %.2e (cond %.2e)" % (name, rel, condC)
    errK = np.max(np.abs(g.Kphi - o.Kphi))
    assert errK <= u * condC * np.abs(o.Cdoubleprime).max(), "Kphi abs err %.2e" % errK
    boundKinv = condK * (errK / np.abs(o.Kphi).max() + u)
    if boundKinv < 1e-2:
        rel = np.max(np.abs(g.Kinv - o.Kinv)) / np.abs(o.Kinv).max()
        assert rel <= 10 * boundKinv, "Kinv rel err %.2e (bound %.2e)" % (rel, boundKinv)


def test_kernel_closed_forms(pkg):
    """test/test_kernels.jl:36,73: k_rbf(0.5, 2.0) and k_matern52(1.0, 1.4)."""
    g = _setup(pkg, "rbf", 2.0, 1.5, np.array([0.5, 2.0]), 1, 1e-6)
    assert np.isclose(g.C[0, 1], 2.0 * np.exp(-(1.5 ** 2) / (2 * 1.5 ** 2)), rtol=1e-13)
    g = _setup(pkg, "matern52", 1.5, 0.8, np.array([1.0, 1.4]), 1, 1e-6)
    r = 0.4 / 0.8
    assert np.isclose(g.C[0, 1], 1.5 * (1 + np.sqrt(5) * r + 5 * r * r / 3) * np.exp(-np.sqrt(5) * r), rtol=1e-13)
    assert np.array_equal(g.C, g.C.T)


def test_fallback_complexity0_and_single_point(pkg):
    """test/test_gp.jl:417-465 (complexity=0) and :467-500 (N=1): C'=C''=m=0, K=eI, Kinv=I/e; 1x1 Cinv = 1/(var+e)."""
    t = np.arange(0.0, 1.0 + 1e-9, 0.25)
    g = _setup(pkg, "matern52", 1.2, 0.7, t, 1, 1e-5, complexity=0)
    n = len(t)
    assert not g.Cprime.any() and not g.Cdoubleprime.any() and not g.mphi.any()
    assert np.allclose(g.Kphi, 1e-5 * np.eye(n), atol=1e-15) and np.allclose(g.Kinv, np.eye(n) / 1e-5, rtol=1e-9)
    g1 = _setup(pkg, "matern52", 1.5, 0.8, np.array([0.0]), 0, 1e-6)
    assert np.isclose(g1.Cinv[0, 0], 1.0 / (1.5 + 1e-6), rtol=1e-12) and g1.mphi[0, 0] == 0.0
    assert np.isclose(g1.Kinv[0, 0], 1e6, rtol=1e-9)
    assert g1.CinvBand.shape == (1, 1)


@pytest.mark.parametrize("n,T,ell", [(201, 20.0, 1.5), (397, 20.0, 1.5), (1281, 64.0, 1.5)])
def test_stable_mode_backward_errors_at_baseline_scale(pkg, n, T, ell):
    """in the BASELINE regime FP64 K+eI from the reference formula is not positive definite, so
    element-wise parity is undefined; the stable route must stay PD (no repaired pivots) and satisfy the
    backward-error identities."""
    t = np.linspace(0.0, T, n)
    g = _setup(pkg, "matern52", 2.0, ell, t, 20, 1e-6, mode="stable")
    I = np.eye(n)
    assert g.repaired_pivots == (0, 0)
    Cj = g.C + 1e-6 * I
    assert np.linalg.norm(Cj @ g.Cinv - I) / (np.linalg.norm(Cj) * np.linalg.norm(g.Cinv)) < 1e-13
    assert np.linalg.norm(g.Kphi @ g.Kinv - I) / (np.linalg.norm(g.Kphi) * np.linalg.norm(g.Kinv)) < 1e-13
    assert np.linalg.eigvalsh(g.Kphi).min() > 0
    assert np.linalg.norm(g.mphi @ Cj - g.Cprime) / (np.linalg.norm(g.mphi) * np.linalg.norm(Cj)) < 1e-13
    if n == 201:
        o = mo.calculate_gp_covariances(mo.MATERN52, [2.0, ell], t, 20, jitter=1e-6, setup_mode="stable", dtype=np.longdouble)
        lam = np.linalg.eigvalsh(g.Kphi).min()
        lam_true = np.linalg.eigvalsh(np.asarray(o.Kphi, dtype=np.float64)).min()
        assert abs(lam - lam_true) < 1e-2 * lam_true
        assert np.max(np.abs(g.Kphi - np.asarray(o.Kphi, dtype=np.float64))) < 1e-8


def test_reference_order_repairs_pivots_at_baseline_scale(pkg):
    """n=201, l=1.5: the reference route gives an indefinite K+eI in FP64; like cholesky(Positive, .) the device
    factorisation must not fail, and reports the repaired pivots."""
    t = np.linspace(0.0, 20.0, 201)
    g = _setup(pkg, "matern52", 2.0, 1.5, t, 20, 1e-6, mode="reference_order")
    assert np.all(np.isfinite(g.Kinv)) and np.all(np.isfinite(g.KinvBand))
    assert g.repaired_pivots[0] == 0


def test_create_with_device_setup_matches_injected_tables(pkg):
    """magi_create running K3-K6 itself == GPCov computed stand-alone and injected (both device routes, all dims batched)."""
    prob = H.make_problem(model="fn", n=41, T=20.0, b=20, n_chains=6, seed=3)
    phi = np.array(prob["phis"]).T
    tg = pkg.MagiTarget.from_config(prob["Y"], prob["tvec"], phi, pkg.fn_system(), prob["sigma_init"], bandsize=20, jitter=1e-6)
    covs = []
    for d in range(2):
        g = pkg.GPCov()
        pkg.calculate_gp_covariances(g, pkg.create_matern52_kernel(*prob["phis"][d]), prob["phis"][d], prob["tvec"], 20, complexity=2, jitter=1e-6)
        covs.append(g)
        assert np.array_equal(tg.get_band_table(d, "KinvBand"), g.KinvBand)
        assert np.array_equal(tg.get_matrix(d, "Kinv"), g.Kinv)
        assert tg.setup_status(d) == (0, 0)
    tg2 = pkg.MagiTarget(prob["Y"], covs, pkg.fn_system(), prob["sigma_init"], [1.0, 1.0, 1.0], 41, 2, 3, False)
    l1, g1 = tg.logdensity_and_gradient_batched(prob["params"])
    l2, g2 = tg2.logdensity_and_gradient_batched(prob["params"])
    assert np.array_equal(l1, l2) and np.array_equal(g1, g2)
    # and the whole thing (setup + evaluation) against the oracle, benign regime
    ll_ref, g_ref = H.oracle_batched(prob)
    rel = np.abs(l1 - ll_ref) / np.abs(ll_ref)
    assert rel.max() < 1e-6


@pytest.mark.parametrize("nu,closed", [(0.5, lambda r: np.exp(-r)), (1.5, lambda r: (1 + np.sqrt(3) * r) * np.exp(-np.sqrt(3) * r)),
                                       (2.5, lambda r: (1 + np.sqrt(5) * r + 5 * r * r / 3) * np.exp(-np.sqrt(5) * r))])
def test_general_matern_kernel_takes_the_zero_derivative_fallback(pkg, nu, closed):
    """src/gaussian_process.jl:278-280: a base kernel other than Matern52Kernel / SqExponentialKernel (here MaternKernel(nu),
    src/kernels.jl:109-118) gets its C, a warning, and C' = C'' = m = 0, K = eI, Kinv = I/e (:319-331) -- not an error."""
    t = np.arange(0.0, 1.0 + 1e-9, 0.25)
    n, var, ell, eps = len(t), 1.3, 0.6, 1e-5
    g = pkg.GPCov()
    pkg.calculate_gp_covariances(g, pkg.create_general_matern_kernel(var, ell, nu), [var, ell], t, 2, complexity=2, jitter=eps)
    r = np.abs(t[:, None] - t[None, :]) / ell
    assert np.allclose(g.C, var * closed(r), rtol=1e-13, atol=0)
    assert not g.Cprime.any() and not g.Cdoubleprime.any() and not g.mphi.any()
    assert np.allclose(g.Kphi, eps * np.eye(n), atol=1e-18) and np.allclose(g.Kinv, np.eye(n) / eps, rtol=1e-9)
    assert np.max(np.abs((g.C + eps * np.eye(n)) @ g.Cinv - np.eye(n))) < 1e-6
    o = mo.calculate_gp_covariances(mo.KERNEL_IDS[g.kernel.kind], [var, ell], t, 2, complexity=2, jitter=eps)
    assert np.max(np.abs(g.Cinv - o.Cinv)) <= 1e-9 * np.abs(o.Cinv).max() and np.array_equal(g.mphiBand, o.mphiBand)


def test_zero_derivative_fallback_is_decided_per_dimension(pkg):
    """src/gaussian_process.jl:299 runs once per GPCov: a dimension whose C' is all zero (here: a lengthscale so short that
    every off-diagonal entry underflows) takes the fallback alone; the other dimension gets the full tables."""
    prob = H.make_problem(model="fn", n=21, T=10.0, b=5, n_chains=4, seed=9)
    phi = np.array([[2.0, 1.0], [1.5, 1e-4]])          # row 0 variances, row 1 lengthscales: dimension 1 underflows
    tg = pkg.MagiTarget.from_config(prob["Y"], prob["tvec"], phi, pkg.fn_system(), prob["sigma_init"], bandsize=5, jitter=1e-6)
    g0 = pkg.GPCov()
    pkg.calculate_gp_covariances(g0, pkg.create_matern52_kernel(2.0, 1.5), [2.0, 1.5], prob["tvec"], 5, complexity=2, jitter=1e-6)
    for name in ("CinvBand", "mphiBand", "KinvBand"):
        assert np.array_equal(tg.get_band_table(0, name), getattr(g0, name)), name
    assert np.any(tg.get_matrix(0, "mphi") != 0.0)
    n = 21
    assert not tg.get_matrix(1, "Cprime").any() and not tg.get_matrix(1, "mphi").any()
    assert np.allclose(tg.get_matrix(1, "Kinv"), np.eye(n) / 1e-6, rtol=1e-9) and np.allclose(tg.get_matrix(1, "Kphi"), 1e-6 * np.eye(n), atol=1e-20)
    ll, g = tg.logdensity_and_gradient_batched(prob["params"])
    assert np.all(np.isfinite(ll)) and np.all(np.isfinite(g))
