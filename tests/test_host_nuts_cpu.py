"""Host logic of run_nuts_sampler (the reference's caller of the single-chain boundary, src/samplers.jl:114-194) on a target
with a known answer: no GPU involved, the target below is a stand-in implementing the same four LogDensityProblems methods."""
import numpy as np

from manifold_constrained_gaussian_process_inference_b200 import samplers as S


class GaussianTarget:
    """log pi(q) = -1/2 sum ((q - mu) / sd)^2"""
    def __init__(self, mu, sd):
        self.mu, self.sd, self.calls = np.asarray(mu, float), np.asarray(sd, float), 0

    def dimension(self):
        return self.mu.shape[0]

    def logdensity(self, q):
        return float(-0.5 * np.sum(((q - self.mu) / self.sd) ** 2))

    def logdensity_and_gradient(self, q):
        self.calls += 1
        if q.shape[0] != self.dimension():
            return -np.inf, np.full(self.dimension(), np.nan)                # interface.jl:179-182
        if not np.all(np.isfinite(q)):
            return -np.inf, np.zeros_like(self.mu)                             # interface.jl:222-226
        return self.logdensity(q), -(q - self.mu) / self.sd ** 2


def test_nuts_recovers_a_gaussian():
    mu, sd = np.array([1.0, -2.0, 0.5, 10.0]), np.array([0.1, 1.0, 3.0, 0.02])
    tg = GaussianTarget(mu, sd)
    chain, stats = S.run_nuts_sampler(tg, np.zeros(4), n_samples=1500, n_adapts=700, initial_step_size=0.05, seed=3)
    assert chain.shape == (800, 4) and len(stats) == 800
    assert np.all(np.abs(chain.mean(axis=0) - mu) < 0.25 * sd)
    assert np.all(np.abs(chain.std(axis=0) / sd - 1.0) < 0.2)
    acc = np.mean([s["accept_stat"] for s in stats])
    assert 0.6 < acc < 0.98 and not any(s["divergent"] for s in stats)
    assert max(s["depth"] for s in stats) <= 10


def test_nuts_asserts_like_the_reference():
    tg = GaussianTarget(np.zeros(3), np.ones(3))
    try:
        S.run_nuts_sampler(tg, np.zeros(4), n_samples=10, n_adapts=5)        # samplers.jl:125
        raise SystemExit("dimension mismatch not detected")
    except AssertionError as e:
        assert "dimension mismatch" in str(e)
    import warnings
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        chain, stats = S.run_nuts_sampler(tg, np.array([0.0, np.inf, 0.0]), n_samples=10, n_adapts=5)   # (-Inf, zeros) aborts the run (:58-60, :186-190)
    assert chain is None and stats is None and any("NUTS" in str(x.message) for x in w)
    v, g = S.logdensity_and_gradient_func_wrapper(tg, np.array([0.1, 0.2, 0.3]))
    assert np.isfinite(v) and g.shape == (3,) and S.logdensity_func_wrapper(tg, np.array([0.1, 0.2, 0.3])) == v


def test_leapfrog_is_reversible_and_conserves_energy():
    tg = GaussianTarget(np.zeros(5), np.array([0.5, 1.0, 2.0, 1.5, 0.7]))
    lpg = lambda q: S.logdensity_and_gradient_func_wrapper(tg, q)
    rng = np.random.default_rng(0)
    q0, p0, minv = rng.normal(size=5), rng.normal(size=5), np.ones(5)
    lp0, g0 = lpg(q0)
    q, p, lp, g = q0, p0, lp0, g0
    for _ in range(50):
        q, p, lp, g = S._leapfrog(lpg, q, p, g, 0.01, minv)
    h0, h1 = -lp0 + 0.5 * p0 @ p0, -lp + 0.5 * p @ p
    assert abs(h1 - h0) < 1e-3
    for _ in range(50):
        q, p, lp, g = S._leapfrog(lpg, q, p, g, -0.01, minv)
    assert np.allclose(q, q0, atol=1e-12) and np.allclose(p, p0, atol=1e-12)
