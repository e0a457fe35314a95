"""Parity at the sizes BASELINE.json quotes: config 3 in dense mode (Lotka-Volterra, n = 1281, band =
n - 1, 2048 chains: the stream-K FP64 DMMA GEMM path with its skinny remainder row) and config 4 (Lorenz-96, D = 64, n = 2001,
band 20, 64 chains: the on-device covariance build / blocked Cholesky / inverse and the banded evaluation of a many-component
model).  The oracle's own setup is O(n^3) numpy per dimension (34 s at n = 2001), so the band tables are the DEVICE's, read
back through magi_get_matrix and handed to the C restatement of the reference's loop (oracle/magi_oracle.c): the 1e-10
claim is made downstream of identical tables, and the device setup is checked separately by its backward
errors.  Tolerance: tests/helpers.py (1e-10 relative)."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import magi_oracle as mo
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _oracle_target_from_device_tables(tg, tvec, phi, yobs, model_id, sigma_init, beta, b):
    covs = []
    for d in range(yobs.shape[1]):
        covs.append(mo.GPCov(phi=np.asarray(phi[:, d]), tvec=tvec, kernel=mo.MATERN52, bandsize=b, CinvBand=tg.get_band_table(d, "CinvBand"),
                             mphiBand=tg.get_band_table(d, "mphiBand"), KinvBand=tg.get_band_table(d, "KinvBand")))
    return mo.make_target(yobs, covs, model_id, sigma_init, beta, False)


def test_config3_dense_lv_n1281_2048_chains(pkg):
    """BASELINE config 3, dense mode: every chain finite, 14 sampled chains against the C oracle at 1e-10, and the whole batch
    against the same chains evaluated in small batches (the non-persistent GEMM)."""
    from manifold_constrained_gaussian_process_inference_b200 import synthetic
    w = synthetic.make_workload("lv1281", 2048)
    n, D = w["n"], w["D"]
    tg = pkg.MagiTarget.from_config(w["yobs"], w["tvec"], w["phi"], pkg.lv_system(), w["sigma_init"], bandsize=n - 1, jitter=1e-6,
                                    setup_mode="stable")
    assert tg.bandsize == n - 1 and all(tg.setup_status(d) == (0, 0) for d in range(D))
    ll, g = tg.logdensity_and_gradient_batched(w["params"])
    assert np.all(np.isfinite(ll)) and np.all(np.isfinite(g))
    tgt = _oracle_target_from_device_tables(tg, w["tvec"], w["phi"], w["yobs"], mo.MODEL_LV, w["sigma_init"], w["beta"], n - 1)
    rng = np.random.default_rng(3)
    idx = np.unique(np.concatenate([[0, 1, 127, 128, 1023, 2046, 2047], rng.integers(0, 2048, size=7)]))
    ll_ref, g_ref = c_oracle.batched(tgt, w["params"][idx])
    H.assert_parity(ll[idx], g[idx], ll_ref, g_ref, "config 3 dense (LV n=1281, b=n-1)")
    sub = np.concatenate([idx, np.arange(300, 340)])
    ll2, g2 = tg.logdensity_and_gradient_batched(w["params"][sub])
    H.assert_parity(ll[sub], g[sub], ll2, g2, "config 3 dense: 2048-chain batch vs small batch")
    # the banded evaluation (b = 20) of the same problem, for the record of config 3's other half
    tb = pkg.MagiTarget.from_config(w["yobs"], w["tvec"], w["phi"], pkg.lv_system(), w["sigma_init"], bandsize=20, jitter=1e-6, setup_mode="stable")
    llb, gb = tb.logdensity_and_gradient_batched(w["params"])
    tgtb = _oracle_target_from_device_tables(tb, w["tvec"], w["phi"], w["yobs"], mo.MODEL_LV, w["sigma_init"], w["beta"], 20)
    llb_ref, gb_ref = c_oracle.batched(tgtb, w["params"][idx])
    H.assert_parity(llb[idx], gb[idx], llb_ref, gb_ref, "config 3 banded (LV n=1281, b=20)")


def _l96_problem():
    rng = np.random.default_rng(20251018 + 3)
    n, D = 2001, 64
    tvec = np.linspace(0.0, 20.0, n)
    phi = np.stack([rng.uniform(10, 20, D), rng.uniform(0.2, 0.4, D)])
    Y = np.full((n, D), np.nan)
    Y[::10] = 8.0 + rng.normal(size=(len(tvec[::10]), D))
    params = np.concatenate([8.0 + rng.normal(size=(64, n * D)), 8.0 + 0.1 * rng.normal(size=(64, 1)),
                             np.log(0.5) + 0.1 * rng.normal(size=(64, D))], axis=1)
    return n, D, tvec, phi, Y, params


def test_config4_lorenz96_d64_n2001(pkg):
    """BASELINE config 4: device setup of 64 distinct dimensions at n = 2001 (backward-error identities, stand-alone setup ==
    setup inside magi_create, band rule) and the 64-chain evaluation against the C oracle."""
    n, D, tvec, phi, Y, params = _l96_problem()
    sig = np.full(D, 0.5)
    tg = pkg.MagiTarget.from_config(Y, tvec, phi, pkg.get_ode_system("lorenz96", D), sig, bandsize=20, jitter=1e-6, setup_mode="stable")
    assert tg.dimension() == n * D + 1 + D == 128129
    assert all(tg.setup_status(d) == (0, 0) for d in range(D))           # stable route: positive definite, no repaired pivots
    I = np.eye(n)
    for d in (0, 31, 63):
        C, Cinv, Cp = tg.get_matrix(d, "C"), tg.get_matrix(d, "Cinv"), tg.get_matrix(d, "Cprime")
        m, K, Kinv = tg.get_matrix(d, "mphi"), tg.get_matrix(d, "Kphi"), tg.get_matrix(d, "Kinv")
        Cj = C + 1e-6 * I
        assert np.allclose(np.diag(C), phi[0, d], rtol=1e-13)                                        # test/test_gp.jl:75
        assert np.linalg.norm(Cj @ Cinv - I) / (np.linalg.norm(Cj) * np.linalg.norm(Cinv)) < 1e-13   # :83 as a backward error
        assert np.linalg.norm(K @ Kinv - I) / (np.linalg.norm(K) * np.linalg.norm(Kinv)) < 1e-13      # :197
        assert np.linalg.norm(m @ Cj - Cp) / (np.linalg.norm(m) * np.linalg.norm(Cj)) < 1e-13         # :162
        assert np.array_equal(Cinv, Cinv.T) and np.array_equal(Kinv, Kinv.T) and np.array_equal(K, K.T)
        for dense, name in ((Cinv, "CinvBand"), (m, "mphiBand"), (Kinv, "KinvBand")):                 # band rule, test/test_gp_utils.jl
            T = tg.get_band_table(d, name)
            for off in (-20, -7, 0, 13, 20):
                ii = np.arange(max(0, -off), min(n, n - off))
                assert np.array_equal(T[20 + off, ii], dense[ii, ii + off])
    for d in (5, 40):                                                     # batched setup inside magi_create == stand-alone setup
        g = pkg.GPCov()
        pkg.calculate_gp_covariances(g, pkg.create_matern52_kernel(phi[0, d], phi[1, d]), phi[:, d], tvec, 20, complexity=2, jitter=1e-6,
                                     setup_mode="stable")
        for name in ("CinvBand", "mphiBand", "KinvBand"):
            assert np.array_equal(tg.get_band_table(d, name), getattr(g, name)), (d, name)
    ll, g = tg.logdensity_and_gradient_batched(params)
    assert np.all(np.isfinite(ll)) and np.all(np.isfinite(g))
    tgt = _oracle_target_from_device_tables(tg, tvec, phi, Y, mo.MODEL_L96, sig, (1.0, 1.0, 1.0), 20)
    idx = np.array([0, 1, 31, 62, 63])
    ll_ref, g_ref = c_oracle.batched(tgt, params[idx])
    H.assert_parity(ll[idx], g[idx], ll_ref, g_ref, "config 4 (Lorenz-96 D=64 n=2001 b=20)")


def test_config5_shape_65536_chains_is_shard_invariant(pkg):
    """BASELINE config 5 at full size on one GPU: 65 536 chains (213 MB of state, larger than L2) in one call; a sample against
    the oracle; a shard of the batch evaluated on its own gives the same bits (chains do not interact)."""
    from manifold_constrained_gaussian_process_inference_b200 import synthetic
    w = synthetic.make_workload("fn201", 65536)
    tg = pkg.MagiTarget.from_config(w["yobs"], w["tvec"], w["phi"], pkg.fn_system(), w["sigma_init"], bandsize=20, jitter=1e-6, setup_mode="stable")
    ll, g = tg.logdensity_and_gradient_batched(w["params"])
    assert np.all(np.isfinite(ll)) and np.all(np.isfinite(g))
    tgt = _oracle_target_from_device_tables(tg, w["tvec"], w["phi"], w["yobs"], mo.MODEL_FN, w["sigma_init"], w["beta"], 20)
    idx = np.unique(np.concatenate([[0, 8191, 8192, 32767, 65535], np.random.default_rng(5).integers(0, 65536, size=11)]))
    ll_ref, g_ref = c_oracle.batched(tgt, w["params"][idx])
    H.assert_parity(ll[idx], g[idx], ll_ref, g_ref, "config 5 (FN n=201, 65 536 chains)")
    lo, hi = 3 * 8192, 4 * 8192                                            # the shard rank 3 of 8 would own
    ll_s, g_s = tg.logdensity_and_gradient_batched(w["params"][lo:hi])
    assert np.array_equal(ll_s, ll[lo:hi]) and np.array_equal(g_s, g[lo:hi])
