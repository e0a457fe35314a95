"""Both K1 kernels against the oracle and against each other.  The library runs small batches (at most one block per SM) on the
dataflow kernel (flow_kernel.cuh; 8 or 16 chains per block) when the state of a block fits shared memory, everything else on
the windowed kernel (banded_kernel.cuh); MAGI_K1 forces one at create time (development knob), so every shape below runs through
both code paths on identical inputs.
Tolerance: tests/helpers.py (1e-10 relative)."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _target(pkg, prob, monkeypatch, variant):
    monkeypatch.setenv("MAGI_K1", variant)
    return H.cuda_target(pkg, prob)


@pytest.mark.parametrize("model,n,b,nc,kw", [
    ("fn", 201, 20, 40, {}),                                 # BASELINE config 2 shape: 3 chain blocks, the last one partial
    ("fn", 201, 20, 16, {"beta": (2.0, 3.0, 5.0)}),
    ("fn", 41, 6, 37, {}),
    ("fn", 16, 0, 8, {}),                                    # diagonal band: NCH = 2, no halo pairs
    ("fn", 9, 1, 3, {}),
    ("fn", 1, 0, 2, {"obs_every": 1}),
    ("fn", 50, 16, 12, {"sigma_fixed": True}),
    ("fn", 64, 32, 17, {}),                                  # widest band of the DMMA tiling (HB = 8)
    ("fn", 120, 9, 33, {"T": 12.0}),
    ("lv", 81, 20, 10, {"T": 4.0}),
    ("hes1", 33, 5, 13, {}),                                 # D = 3
    ("hes1", 64, 12, 20, {}),
    ("fn", 397, 20, 11, {"beta": (1.0, 1.0, 5.0), "obs_every": 4, "T": 20.0}),   # BASELINE config 1 shape: 8 chains per block only
])
def test_flow_and_windowed_kernels_match_the_oracle(pkg, monkeypatch, model, n, b, nc, kw):
    prob = H.make_problem(model=model, n=n, b=b, n_chains=nc, seed=7 * n + b, **kw)
    ll_ref, g_ref = H.oracle_batched(prob)
    out = {}
    for variant in ("flow", "windowed"):
        tg = _target(pkg, prob, monkeypatch, variant)
        ll, g = tg.logdensity_and_gradient_batched(prob["params"])
        H.assert_parity(ll, g, ll_ref, g_ref, "%s %s n=%d b=%d" % (variant, model, n, b))
        ll2, _ = tg.logdensity_and_gradient_batched(prob["params"], want_grad=False)
        assert np.array_equal(ll, ll2)
        out[variant] = (ll, g)
        tg.close()
    H.assert_parity(out["flow"][0], out["flow"][1], out["windowed"][0], out["windowed"][1], "flow vs windowed")


def test_flow_kernel_is_persistent_and_deterministic(pkg, monkeypatch):
    """More chain blocks than SMs (every block of the persistent grid handles several; the mbarrier phases alternate), a
    partial last block, and bit-identical results from two launches (units are drawn dynamically, sums are not)."""
    nc = 148 * 16 * 2 + 16 * 5 + 3
    base = H.make_problem(model="fn", n=201, b=20, n_chains=8, seed=11, T=20.0, obs_every=5)
    rng = np.random.default_rng(3)
    params = np.repeat(base["params"], (nc + 7) // 8, axis=0)[:nc] + 1e-3 * rng.normal(size=(nc, base["params"].shape[1]))
    tg = _target(pkg, base, monkeypatch, "flow")
    ll, g = tg.logdensity_and_gradient_batched(params)
    ll_b, g_b = tg.logdensity_and_gradient_batched(params)
    assert np.array_equal(ll, ll_b) and np.array_equal(g, g_b)
    idx = np.unique(np.concatenate([[0, 1, 15, 16, nc // 2, nc - 4, nc - 3, nc - 2, nc - 1], rng.integers(0, nc, size=9)]))
    ll_ref, g_ref = H.oracle_batched(base, params[idx])
    H.assert_parity(ll[idx], g[idx], ll_ref, g_ref, "flow kernel, persistent grid")
    # a chain gives the same bits wherever it sits in the batch
    sub = np.concatenate([idx, np.arange(200, 230)])
    ll2, g2 = tg.logdensity_and_gradient_batched(params[sub])
    assert np.array_equal(ll[sub], ll2) and np.array_equal(g[sub], g2)


def test_flow_kernel_guards_are_per_chain(pkg, monkeypatch):
    """interface.jl:222-226, 260-264 per chain: a poisoned chain returns (-Inf, zeros) and leaves its neighbours alone."""
    prob = H.make_problem(model="fn", n=41, b=6, n_chains=20, seed=5)
    tg = _target(pkg, prob, monkeypatch, "flow")
    ll0, g0 = tg.logdensity_and_gradient_batched(prob["params"])
    bad = prob["params"].copy()
    bad[3, 7] = np.nan                   # a state value
    bad[9, 2 * 41 + 2] = np.inf          # theta_3
    bad[17, 2 * 41 + 3] = np.nan         # log sigma_1
    ll, g = tg.logdensity_and_gradient_batched(bad)
    for c in (3, 9, 17):
        assert ll[c] == -np.inf and np.all(g[c] == 0.0)
    keep = np.setdiff1d(np.arange(20), [3, 9, 17])
    assert np.array_equal(ll[keep], ll0[keep]) and np.array_equal(g[keep], g0[keep])
    ll_ref, g_ref = H.oracle_batched(prob, bad)
    H.assert_parity(ll, g, ll_ref, g_ref, "guards")


@pytest.mark.parametrize("nc", [1, 8 * 148, 8 * 148 + 1, 16 * 148, 16 * 148 + 1])
def test_dispatch_by_batch_size_is_seamless(pkg, monkeypatch, nc):
    """Default dispatch: <= 8 chains per SM -> dataflow kernel with 8 chains per block, <= 16 per SM -> 16 chains per block,
    beyond -> windowed kernel.  Same chains, same values to rounding on either side of every switch; a sample against the oracle."""
    monkeypatch.delenv("MAGI_K1", raising=False)
    base = H.make_problem(model="fn", n=201, b=20, n_chains=8, seed=17, T=20.0, obs_every=5)
    rng = np.random.default_rng(nc)
    params = np.repeat(base["params"], (nc + 7) // 8, axis=0)[:nc] + 1e-3 * rng.normal(size=(nc, base["params"].shape[1]))
    tg = H.cuda_target(pkg, base)
    ll, g = tg.logdensity_and_gradient_batched(params)
    idx = np.unique(np.concatenate([[0, nc // 2, nc - 1], rng.integers(0, nc, size=5)]))
    ll_ref, g_ref = H.oracle_batched(base, params[idx])
    H.assert_parity(ll[idx], g[idx], ll_ref, g_ref, "default dispatch, %d chains" % nc)
    monkeypatch.setenv("MAGI_K1", "windowed")
    tw = H.cuda_target(pkg, base)
    llw, gw = tw.logdensity_and_gradient_batched(params)
    H.assert_parity(ll, g, llw, gw, "default dispatch vs windowed, %d chains" % nc)



@pytest.mark.parametrize("model,n,b,nc,kw", [
    ("fn", 201, 4, 300, {"T": 20.0, "obs_every": 5}),        # three blocks of 128 chains, the last one partial
    ("fn", 201, 2, 140, {"beta": (2.0, 3.0, 5.0)}),
    ("fn", 201, 3, 129, {}),
    ("fn", 41, 1, 37, {}),
    ("fn", 16, 0, 8, {}),                                    # diagonal band
    ("fn", 9, 1, 3, {}),
    ("fn", 7, 4, 5, {}),                                     # band wider than half the time axis
    ("fn", 1, 0, 2, {"obs_every": 1}),
    ("fn", 50, 4, 12, {"sigma_fixed": True}),
    ("lv", 81, 3, 130, {"T": 4.0}),
    ("lv", 33, 2, 17, {"T": 4.0, "beta": (1.0, 1.0, 5.0)}),
])
def test_narrow_kernel_matches_the_oracle(pkg, monkeypatch, model, n, b, nc, kw):
    """K1-narrow (narrow_kernel.cuh: FP64-FMA sweep, one thread per chain, band half-widths <= 4) against the oracle and against
    the windowed DMMA kernel on identical inputs."""
    prob = H.make_problem(model=model, n=n, b=b, n_chains=nc, seed=11 * n + b, **kw)
    ll_ref, g_ref = H.oracle_batched(prob)
    tg = _target(pkg, prob, monkeypatch, "narrow")
    ll, g = tg.logdensity_and_gradient_batched(prob["params"])
    H.assert_parity(ll, g, ll_ref, g_ref, "narrow %s n=%d b=%d" % (model, n, b))
    ll2, _ = tg.logdensity_and_gradient_batched(prob["params"], want_grad=False)
    assert np.array_equal(ll, ll2)
    ll3, g3 = tg.logdensity_and_gradient_batched(prob["params"][::-1].copy())      # a chain gives the same bits wherever it sits
    assert np.array_equal(ll, ll3[::-1]) and np.array_equal(g, g3[::-1])
    tg.close()
    tw = _target(pkg, prob, monkeypatch, "windowed")
    llw, gw = tw.logdensity_and_gradient_batched(prob["params"])
    H.assert_parity(ll, g, llw, gw, "narrow vs windowed")


def test_narrow_kernel_guards_are_per_chain(pkg, monkeypatch):
    """interface.jl:222-226, 260-264 per chain on K1-narrow."""
    prob = H.make_problem(model="fn", n=41, b=3, n_chains=20, seed=5)
    tg = _target(pkg, prob, monkeypatch, "narrow")
    ll0, g0 = tg.logdensity_and_gradient_batched(prob["params"])
    bad = prob["params"].copy()
    bad[3, 7] = np.nan
    bad[9, 2 * 41 + 2] = np.inf
    bad[17, 2 * 41 + 3] = np.nan
    ll, g = tg.logdensity_and_gradient_batched(bad)
    for c in (3, 9, 17):
        assert ll[c] == -np.inf and np.all(g[c] == 0.0)
    keep = np.setdiff1d(np.arange(20), [3, 9, 17])
    assert np.array_equal(ll[keep], ll0[keep]) and np.array_equal(g[keep], g0[keep])
    ll_ref, g_ref = H.oracle_batched(prob, bad)
    H.assert_parity(ll, g, ll_ref, g_ref, "guards")


@pytest.mark.parametrize("b,nc", [(2, 32 * 148), (2, 32 * 148 + 1), (4, 64 * 148), (4, 64 * 148 + 1), (2, 96 * 148 + 5), (1, 40000), (4, 257 * 148)])
def test_narrow_dispatch_is_seamless(pkg, monkeypatch, b, nc):
    """Default dispatch for band half-widths <= 4: large batches run on K1-narrow (beyond 32 chains per SM for b <= 2, 64 per SM for b = 3, 4; its two organisations of the copies --
    warp-specialised up to 96 chains per SM, every thread copying beyond -- on either side of their own seams),
    through the chunked host path as well; same values to rounding as the windowed kernel, a sample against the oracle."""
    monkeypatch.delenv("MAGI_K1", raising=False)
    base = H.make_problem(model="fn", n=201, b=b, n_chains=8, seed=23, T=20.0, obs_every=5)
    rng = np.random.default_rng(nc + b)
    params = np.repeat(base["params"], (nc + 7) // 8, axis=0)[:nc] + 1e-3 * rng.normal(size=(nc, base["params"].shape[1]))
    tg = H.cuda_target(pkg, base)
    ll, g = tg.logdensity_and_gradient_batched(params)
    idx = np.unique(np.concatenate([[0, nc // 2, nc - 1], rng.integers(0, nc, size=5)]))
    ll_ref, g_ref = H.oracle_batched(base, params[idx])
    H.assert_parity(ll[idx], g[idx], ll_ref, g_ref, "default dispatch, b=%d, %d chains" % (b, nc))
    monkeypatch.setenv("MAGI_K1", "windowed")
    tw = H.cuda_target(pkg, base)
    llw, gw = tw.logdensity_and_gradient_batched(params)
    H.assert_parity(ll, g, llw, gw, "default dispatch vs windowed, b=%d, %d chains" % (b, nc))
