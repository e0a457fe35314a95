"""Parity of the CUDA path (through the C ABI) against the oracle, on identical inputs including the band tables.
Tolerance: 1e-10 relative (north_star), written in tests/helpers.py."""
import numpy as np
import pytest

from oracle import magi_oracle as mo
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _ref_fn_case(missing=False):
    """test/test_likelihoods.jl:18-59 (FN, N=3, RBF σ²=1.5 ℓ=1.2, b=1, ε=1e-5)."""
    t = np.array([0.0, 1.0, 2.0])
    covs = [mo.calculate_gp_covariances(mo.RBF, [1.5, 1.2], t, 1, complexity=2, jitter=1e-5) for _ in range(2)]
    X = np.array([[1.0, 0.5], [1.1, 0.6], [1.2, 0.7]])
    Y = X + np.array([[0.05, -0.02], [-0.01, 0.03], [0.02, 0.01]])
    if missing:
        Y[1, 0] = np.nan
    sig = np.array([0.1, 0.15])
    tgt = mo.make_target(Y, covs, mo.MODEL_FN, sig, (1.0, 1.0, 1.0), True)
    params = np.concatenate([X.reshape(-1, order="F"), [0.5, 0.6, 0.7]])[None, :]
    return dict(target=tgt, params=params, covs=covs)


def test_reference_fn_case_fixed_sigma(pkg):
    prob = _ref_fn_case()
    tg = H.cuda_target(pkg, prob)
    ll, g = tg.logdensity_and_gradient(prob["params"][0])
    ll_ref, g_ref = mo.logdensity_and_gradient(prob["target"], prob["params"][0])
    H.assert_parity([ll], g, [ll_ref], g_ref, "FN N=3")
    ll_gold, g_gold, rtol = H.fn_n3_golden()              # frozen long-double evaluation (tests/golden/reference_known_answers.json)
    assert abs(ll - ll_gold) <= 10 * rtol * abs(ll_gold)
    assert np.max(np.abs(g - g_gold[:9]) / np.maximum(1.0, np.abs(g_gold[:9]))) <= 10 * rtol
    assert tg.dimension() == 9 and tg.capabilities() == pkg.LogDensityOrder(1)
    assert abs(tg.logdensity(prob["params"][0]) - ll) <= 1e-12 * abs(ll)


def test_reference_missing_observation_known_answer(pkg):
    """test/test_likelihoods.jl:106-148: gradient element of the missing observation moves by exactly +1.0."""
    full, miss = _ref_fn_case(False), _ref_fn_case(True)
    ll_f, g_f = H.cuda_target(pkg, full).logdensity_and_gradient(full["params"][0])
    ll_m, g_m = H.cuda_target(pkg, miss).logdensity_and_gradient(miss["params"][0])
    assert ll_m < ll_f
    assert abs((g_m[1] - g_f[1]) - 1.0) < 1e-6
    others = np.delete(np.arange(9), 1)
    assert np.max(np.abs(g_m[others] - g_f[others])) < 1e-6


def test_reference_hes1_case(pkg):
    """test/test_likelihoods.jl:165-179 (Hes1, D=3, k=7)."""
    t = np.array([0.0, 1.0, 2.0])
    covs = [mo.calculate_gp_covariances(mo.RBF, [1.5, 1.2], t, 1, complexity=2, jitter=1e-5) for _ in range(3)]
    X = np.array([[1.0, 2.0, 3.0], [1.1, 2.1, 2.9], [1.2, 2.2, 2.8]])
    Y = X + np.array([[0.01, 0.02, 0.03], [-0.02, -0.01, -0.03], [0.03, 0.01, 0.02]])
    tgt = mo.make_target(Y, covs, mo.MODEL_HES1, [0.1, 0.2, 0.3], (1.0, 1.0, 1.0), False)
    params = np.concatenate([X.reshape(-1, order="F"), [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7], np.log([0.1, 0.2, 0.3])])
    prob = dict(target=tgt, params=params[None, :], covs=covs)
    ll, g = H.cuda_target(pkg, prob).logdensity_and_gradient(params)
    ll_ref, g_ref = mo.logdensity_and_gradient(tgt, params)
    H.assert_parity([ll], g, [ll_ref], g_ref, "Hes1 N=3")


@pytest.mark.parametrize("model,n,b,nc,kw", [
    ("fn", 41, 6, 37, {}),
    ("fn", 41, 40, 9, {}),                                  # full band == dense (test/test_gp.jl:551-585)
    ("fn", 201, 20, 40, {}),                                # BASELINE config 2 shape
    ("fn", 201, 20, 11, {"beta": (2.0, 3.0, 5.0)}),
    ("fn", 397, 20, 8, {"beta": (1.0, 1.0, 5.0), "obs_every": 4, "T": 20.0}),   # BASELINE config 1 shape
    ("fn", 16, 0, 8, {}),
    ("fn", 9, 1, 3, {"kernel": mo.RBF}),
    ("fn", 1, 0, 2, {"obs_every": 1}),
    ("fn", 50, 16, 12, {"sigma_fixed": True}),
    ("fn", 64, 32, 16, {}),
    ("lv", 81, 20, 10, {"T": 4.0}),
    ("hes1", 33, 5, 13, {}),
    ("fn", 120, 119, 7, {"T": 12.0}),                       # dense path: band = n-1 (GEMM formulation)
    ("lv", 150, 40, 5, {"T": 6.0}),                         # dense path with a partial band (half-width > 32)
    ("hes1", 70, 69, 4, {}),
])
def test_batched_parity_random(pkg, model, n, b, nc, kw):
    prob = H.make_problem(model=model, n=n, b=b, n_chains=nc, seed=n + b, **kw)
    tg = H.cuda_target(pkg, prob)
    ll, g = tg.logdensity_and_gradient_batched(prob["params"])
    ll_ref, g_ref = H.oracle_batched(prob)
    H.assert_parity(ll, g, ll_ref, g_ref, "%s n=%d b=%d" % (model, n, b))
    # value-only variant and batched == single-chain
    ll2, _ = tg.logdensity_and_gradient_batched(prob["params"], want_grad=False)
    assert np.array_equal(ll, ll2)
    l1, g1 = tg.logdensity_and_gradient(prob["params"][nc // 2])
    assert l1 == ll[nc // 2] and np.array_equal(g1, g[nc // 2])


@pytest.mark.parametrize("model,n,b,nc,what", [
    ("lv", 264, 263, 9600, "stream-K GEMM, 2 tile rows + 8 remainder rows (skinny kernel), 75 tile columns"),
    ("fn", 300, 299, 3300, "stream-K GEMM with predicated edge tiles in both directions"),
    ("fn", 201, 20, 2368, "K1 with two chain-groups per block, last batch size before the switch"),
    ("fn", 201, 20, 2369, "K1 with four chain-groups per block, partial last block"),
    ("fn", 201, 20, 5000, "K1 pipelined host call: chunks of different block shapes"),
    ("lv", 1281, 20, 2400, "BASELINE config 3 time axis: Ke scratch in L2, four chain-groups per block"),
    ("lv", 1281, 20, 300, "BASELINE config 3 time axis: Ke scratch in L2, two chain-groups per block"),
])
def test_large_batch_paths(pkg, model, n, b, nc, what):
    """Code paths that only large batches reach (persistent stream-K GEMM of the dense mode; the per-call block shape of the
    banded kernel; the chunked host call): a sample of chains against the oracle, every chain against a small-batch call."""
    base = H.make_problem(model=model, n=n, b=b, n_chains=8, seed=n + nc, T=0.06 * n)
    rng = np.random.default_rng(nc)
    params = np.repeat(base["params"], (nc + 7) // 8, axis=0)[:nc] + 1e-3 * rng.normal(size=(nc, base["params"].shape[1]))
    tg = H.cuda_target(pkg, base)
    ll, g = tg.logdensity_and_gradient_batched(params)
    assert np.all(np.isfinite(ll)) and np.all(np.isfinite(g))
    idx = np.unique(np.concatenate([[0, 1, nc // 2, nc - 2, nc - 1], rng.integers(0, nc, size=7)]))
    ll_ref, g_ref = H.oracle_batched(base, params[idx])
    H.assert_parity(ll[idx], g[idx], ll_ref, g_ref, what)
    # the same chains in a small batch (other block shape / the non-persistent GEMM): same values to rounding
    sub = np.concatenate([idx, np.arange(100, 140)])
    ll2, g2 = tg.logdensity_and_gradient_batched(params[sub])
    H.assert_parity(ll[sub], g[sub], ll2, g2, what + " (vs small batch)")


@pytest.mark.parametrize("model,n,b,nc", [("fn", 201, 20, 37), ("fn", 41, 6, 2400), ("lv", 150, 40, 37), ("fn", 120, 119, 200)])
def test_device_call_stays_inside_its_buffers(pkg, model, n, b, nc):
    """The device-pointer entry (magi_logdensity_and_gradient_batched_dev) writes ll[0 : n_chains] and grad[0 : n_chains * P] and
    nothing else: guard bands around both buffers keep their sentinel, for partial blocks (chain counts that are not a multiple
    of 8 / 16 / 32), both block shapes and the GEMM path with its remainder strips."""
    import torch
    prob = H.make_problem(model=model, n=n, b=b, n_chains=8, seed=n + nc, T=0.06 * n)
    rng = np.random.default_rng(nc)
    params = np.repeat(prob["params"], (nc + 7) // 8, axis=0)[:nc] + 1e-3 * rng.normal(size=(nc, prob["params"].shape[1]))
    tg = H.cuda_target(pkg, prob)
    P, guard, sentinel = params.shape[1], 4096, -7.25
    dev = torch.device("cuda")
    p = torch.from_numpy(params).to(dev)
    gbuf = torch.full((nc * P + 2 * guard,), sentinel, dtype=torch.float64, device=dev)
    lbuf = torch.full((nc + 2 * guard,), sentinel, dtype=torch.float64, device=dev)
    tg.logdensity_and_gradient_batched_dev(nc, p.data_ptr(), lbuf.data_ptr() + 8 * guard, gbuf.data_ptr() + 8 * guard, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    g, l = gbuf.cpu().numpy(), lbuf.cpu().numpy()
    assert np.all(g[:guard] == sentinel) and np.all(g[-guard:] == sentinel) and np.all(l[:guard] == sentinel) and np.all(l[-guard:] == sentinel)
    assert np.array_equal(p.cpu().numpy(), params)                      # the inputs are read-only
    ll, grad = tg.logdensity_and_gradient_batched(params)
    assert np.array_equal(l[guard:-guard], ll) and np.array_equal(g[guard:-guard].reshape(nc, P), grad)


def test_guards_per_chain(pkg):
    """interface.jl:179-182, 222-226: wrong length -> (-Inf, NaN...); a non-finite chain -> (-Inf, 0...) without
    poisoning its neighbours."""
    prob = H.make_problem(n=41, b=6, n_chains=10, seed=5)
    tg = H.cuda_target(pkg, prob)
    params = prob["params"].copy()
    params[3, 7] = np.nan
    params[6, -1] = np.inf            # log sigma = +Inf -> clamped to 15 (finite result)
    params[8, 41 * 2] = 0.0           # theta_a = 0 is fine; theta_c = 0 -> division by zero
    params[8, 41 * 2 + 2] = 0.0
    ll, g = tg.logdensity_and_gradient_batched(params)
    ll_ref, g_ref = H.oracle_batched(prob, params)
    assert ll[3] == -np.inf and np.all(g[3] == 0.0)
    assert ll[8] == -np.inf and np.all(g[8] == 0.0)
    H.assert_parity(ll, g, ll_ref, g_ref, "guards")
    l, gg = tg.logdensity_and_gradient(params[0][:-1])
    assert l == -np.inf and np.all(np.isnan(gg)) and gg.shape[0] == tg.dimension()
    assert tg.logdensity(params[0][:-1]) == -np.inf


def test_extreme_theta_and_sparse_obs(pkg):
    """test/test_likelihoods.jl:181-205."""
    prob = _ref_fn_case()
    p = prob["params"][0].copy()
    p[6:9] = [1e-8, 1e8, 1.0]
    ll, g = H.cuda_target(pkg, prob).logdensity_and_gradient(p)
    ll_ref, g_ref = mo.logdensity_and_gradient(prob["target"], p)
    assert np.isfinite(ll) and np.all(np.isfinite(g))
    H.assert_parity([ll], g, [ll_ref], g_ref, "extreme theta")
    Y = np.full((3, 2), np.nan)
    Y[0, 0] = prob["target"].yobs[0, 0]
    Y[2, 1] = prob["target"].yobs[2, 1]
    prob["target"].yobs = Y
    ll, g = H.cuda_target(pkg, prob).logdensity_and_gradient(prob["params"][0])
    ll_ref, g_ref = mo.logdensity_and_gradient(prob["target"], prob["params"][0])
    H.assert_parity([ll], g, [ll_ref], g_ref, "sparse obs")


def test_tempering_changes_result(pkg):
    """test/test_likelihoods.jl:158-163."""
    p0 = H.make_problem(n=21, b=4, n_chains=2, seed=1)
    p1 = H.make_problem(n=21, b=4, n_chains=2, seed=1, beta=(10.0, 1.0, 1.0))
    l0, g0 = H.cuda_target(pkg, p0).logdensity_and_gradient_batched(p0["params"])
    l1, g1 = H.cuda_target(pkg, p1).logdensity_and_gradient_batched(p1["params"])
    assert np.all(l0 != l1) and not np.allclose(g0, g1, atol=1e-6, rtol=1e-6)


def test_lorenz96_dense_path(pkg):
    """Lorenz-96 (D = 16 here; BASELINE config 4 uses 64) is not in the reference: oracle = derivation, FD-checked in
    tests/test_oracle_pins.py.  Runs on the GEMM path with band-truncated operators."""
    rng = np.random.default_rng(9)
    n, D, b, nc = 48, 16, 10, 6
    t = np.linspace(0.0, 2.0, n)
    covs = [mo.calculate_gp_covariances(mo.MATERN52, [10.0 + d, 0.3 + 0.01 * d], t, b, jitter=1e-6) for d in range(D)]
    Y = np.full((n, D), np.nan)
    Y[::4] = 8.0 + rng.normal(size=(len(t[::4]), D))
    tgt = mo.make_target(Y, covs, mo.MODEL_L96, np.full(D, 0.5), (1.0, 1.0, 1.0), False)
    P = mo.dimension(tgt)
    params = np.concatenate([8.0 + rng.normal(size=(nc, n * D)), 8.0 + 0.1 * rng.normal(size=(nc, 1)), np.log(0.5) + 0.1 * rng.normal(size=(nc, D))], axis=1)
    prob = dict(target=tgt, params=params, covs=covs)
    tg = H.cuda_target(pkg, prob)
    ll, g = tg.logdensity_and_gradient_batched(params)
    ll_ref, g_ref = H.oracle_batched(prob)
    H.assert_parity(ll, g, ll_ref, g_ref, "lorenz96")


def test_pipelined_host_call_matches_single_stream(pkg, monkeypatch):
    """Batches >= 2048 chains go through the chunked H2D / kernel / D2H pipeline: results must be bit-identical to
    evaluating the same chains in small calls (with the kernel variant pinned: the library picks it by the size of the call)."""
    monkeypatch.setenv("MAGI_K1", "windowed")
    prob = H.make_problem(n=41, T=8.0, b=6, n_chains=8, seed=21)
    tg = H.cuda_target(pkg, prob)
    rng = np.random.default_rng(0)
    params = np.repeat(prob["params"], 400, axis=0) + 1e-3 * rng.normal(size=(3200, prob["params"].shape[1]))
    ll, g = tg.logdensity_and_gradient_batched(params)
    ll2 = np.concatenate([tg.logdensity_and_gradient_batched(params[i:i + 800])[0] for i in range(0, 3200, 800)])
    g2 = np.concatenate([tg.logdensity_and_gradient_batched(params[i:i + 800])[1] for i in range(0, 3200, 800)])
    assert np.array_equal(ll, ll2) and np.array_equal(g, g2)
    ll_ref, g_ref = H.oracle_batched(prob, params[::457])
    H.assert_parity(ll[::457], g[::457], ll_ref, g_ref, "pipelined")


@pytest.mark.parametrize("name,mid,D,k", [("hes1log", mo.MODEL_HES1LOG, 3, 7), ("hes1log_fixg", mo.MODEL_HES1LOG_FIXG, 3, 6),
                                          ("hes1log_fixf", mo.MODEL_HES1LOG_FIXF, 3, 6), ("hiv", mo.MODEL_HIV, 4, 9), ("ptrans", mo.MODEL_PTRANS, 5, 6)])
def test_models_without_reference_jacobians(pkg, name, mid, D, k):
    """Hes1-log (3 variants), HIV and protein-transduction have a right-hand side in the reference (src/ode_models.jl:83-233)
    but no Jacobians: the log density must match the oracle, and the device gradient (derived Jacobians) must match
    central finite differences of the device log density."""
    rng = np.random.default_rng(40 + mid)
    n, b, nc = 24, 5, 3
    t = np.linspace(0.0, 4.0, n)
    covs = [mo.calculate_gp_covariances(mo.MATERN52, [1.0 + 0.2 * d, 1.0 + 0.1 * d], t, b, jitter=1e-6) for d in range(D)]
    base = 0.3 * np.sin(t[:, None] + np.arange(D)[None, :]) + (1.0 if name == "ptrans" else 0.2)
    Y = np.full((n, D), np.nan); Y[::2] = base[::2] + 0.05 * rng.normal(size=base[::2].shape)
    tgt = mo.make_target(Y, covs, mid, np.full(D, 0.1), (1.0, 1.0, 1.0), False)
    th0 = (np.array([10.0, 1.0, 2.0, 3.0, 4.0, 5.0, 1.0, 1.0, 1.0]) if name == "hiv" else 0.3 + 0.1 * np.arange(k))
    params = np.stack([np.concatenate([(base + 0.05 * rng.normal(size=base.shape)).reshape(-1, order="F"), th0 * np.exp(0.05 * rng.normal(size=k)),
                                       np.log(0.1) + 0.05 * rng.normal(size=D)]) for _ in range(nc)])
    prob = dict(target=tgt, params=params, covs=covs)
    tg = H.cuda_target(pkg, prob)
    ll, g = tg.logdensity_and_gradient_batched(params)
    ll_ref = np.array([mo.logdensity(tgt, p) for p in params])
    assert np.all(np.isfinite(ll)) and np.all(np.isfinite(g))
    assert np.max(np.abs(ll - ll_ref) / np.abs(ll_ref)) <= H.LL_RTOL
    P = params.shape[1]
    idx = rng.choice(P, size=24, replace=False)
    idx = np.unique(np.concatenate([idx, np.arange(n * D, P)]))             # always include theta and log sigma
    p0 = params[0]
    pert = []
    hs = []
    for i in idx:
        h = 1e-5 * max(1.0, abs(p0[i]))
        for sgn in (-2, -1, 1, 2):
            q = p0.copy(); q[i] += sgn * h; pert.append(q)
        hs.append(h)
    lls, _ = tg.logdensity_and_gradient_batched(np.array(pert), want_grad=False)
    lls = lls.reshape(len(idx), 4)
    fd = (lls[:, 0] - 8 * lls[:, 1] + 8 * lls[:, 2] - lls[:, 3]) / (12 * np.array(hs))
    scale = np.maximum(np.abs(fd), 1e-3 * np.abs(g[0]).max())
    assert np.max(np.abs(fd - g[0][idx]) / scale) < 1e-5, (name, np.max(np.abs(fd - g[0][idx]) / scale))
