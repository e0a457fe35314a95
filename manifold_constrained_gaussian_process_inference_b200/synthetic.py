"""Seeded synthetic workloads of SURVEY.md section 8(d) / BASELINE.json `configs` (host-side data generation only)."""
from __future__ import annotations

import numpy as np


def _rk4(f, x0, tvec, substeps=20):
    x = np.array(x0, dtype=np.float64)
    out = np.zeros((len(tvec), len(x0)))
    out[0] = x
    for i in range(1, len(tvec)):
        h = (tvec[i] - tvec[i - 1]) / substeps
        for _ in range(substeps):
            k1 = f(x); k2 = f(x + 0.5 * h * k1); k3 = f(x + 0.5 * h * k2); k4 = f(x + h * k3)
            x = x + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
        out[i] = x
    return out


def fn_rhs(theta):
    a, b, c = theta
    return lambda u: np.array([c * (u[0] - u[0] ** 3 / 3.0 + u[1]), -(u[0] - a + b * u[1]) / c])


def lv_rhs(theta):
    al, be, de, ga = theta
    return lambda u: np.array([al * u[0] - be * u[0] * u[1], de * u[0] * u[1] - ga * u[1]])


CONFIGS = {
    # BASELINE.json configs[1]: FitzHugh-Nagumo synthetic, n=201, 4096 chains, banded K^-1
    "fn201": dict(model="fn", n=201, T=20.0, obs_every=5, noise=0.2, theta=(0.2, 0.2, 3.0), x0=(-1.0, 1.0),
                  phi=((2.0, 1.5), (1.0, 2.0)), bandsize=20, beta=(1.0, 1.0, 1.0), chains=4096, seed=20251018 + 1),
    # configs[0] shape: run_scripts/fn_example.jl (n=397 after fill-level 2, beta=[1,1,5])
    "fn397": dict(model="fn", n=397, T=20.0, obs_every=4, noise=0.2, theta=(0.2, 0.2, 3.0), x0=(-1.0, 1.0),
                  phi=((2.0, 1.5), (1.0, 2.0)), bandsize=20, beta=(1.0, 1.0, 5.0), chains=1, seed=20251018 + 0),
    # configs[2]: Lotka-Volterra, n=1281, 2048 chains (banded b=20 or dense b=n-1)
    "lv1281": dict(model="lv", n=1281, T=64.0, obs_every=16, noise=0.1, theta=(1.5, 1.0, 3.0, 1.0), x0=(1.0, 1.0),
                   phi=((1.0, 1.5), (1.0, 1.5)), bandsize=20, beta=(1.0, 1.0, 1.0), chains=2048, seed=20251018 + 2),
}


def make_workload(name: str, n_chains: int | None = None, rank: int = 0, bandsize: int | None = None):
    """Returns dict(tvec, yobs, phi (2 x D), sigma_init, beta, params (n_chains x P), model, bandsize).  Chain states:
    X = truth + 0.1 N(0,1), theta = theta* exp(0.1 N(0,1)), log sigma = log(noise) + 0.1 N(0,1); per-rank streams are
    keyed by (seed, rank) so a sharded run draws distinct chains on every GPU."""
    c = dict(CONFIGS[name])
    nch = int(n_chains or c["chains"])
    n = c["n"]
    tvec = np.linspace(0.0, c["T"], n)
    rhs = fn_rhs(c["theta"]) if c["model"] == "fn" else lv_rhs(c["theta"])
    truth = _rk4(rhs, c["x0"], tvec)
    D = truth.shape[1]
    rng = np.random.default_rng(c["seed"])
    Y = np.full((n, D), np.nan)
    Y[::c["obs_every"]] = truth[::c["obs_every"]] + c["noise"] * rng.normal(size=truth[::c["obs_every"]].shape)
    crng = np.random.default_rng([c["seed"], rank])
    th = np.asarray(c["theta"])
    X = truth[None] + 0.1 * crng.normal(size=(nch, n, D))
    params = np.concatenate([X.transpose(0, 2, 1).reshape(nch, n * D),          # vec(X): time fastest, then dimension
                             th[None] * np.exp(0.1 * crng.normal(size=(nch, len(th)))),
                             np.log(c["noise"]) + 0.1 * crng.normal(size=(nch, D))], axis=1)
    return dict(name=name, model=c["model"], tvec=tvec, yobs=Y, phi=np.asarray(c["phi"]).T.copy(), sigma_init=np.full(D, c["noise"]),
                beta=c["beta"], params=np.ascontiguousarray(params), bandsize=int(bandsize if bandsize is not None else c["bandsize"]),
                n=n, D=D, k=len(th), truth=truth, theta_true=th)


def algorithmic_bytes_per_eval(P: int) -> int:
    """read P parameters, write P gradient entries + 1 log density."""
    return 8 * (2 * P + 1)


def algorithmic_flops_per_eval(n: int, D: int, b: int, c_model: int = 50) -> int:
    """four band products per dimension (2 flop per stored band entry) + ODE/pointwise work."""
    nnz = n * (2 * b + 1) - b * (b + 1) if b < n - 1 else n * n
    return 4 * D * 2 * nnz + n * (6 * D + c_model)
