"""Independent CPU reference sampler (test infrastructure): vectorised static-trajectory HMC over a few chains whose
gradient is the oracle's C restatement.  Used to check that the on-device sampler targets the same posterior."""
import numpy as np

from oracle import c_oracle


def cpu_hmc(target, params0, n_iter, n_warm, eps, n_leapfrog, minv=None, seed=0):
    rng = np.random.default_rng(seed)
    q = np.array(params0, dtype=np.float64)
    nc, P = q.shape
    minv = np.ones(P) if minv is None else np.asarray(minv)
    ll, g = c_oracle.batched(target, q)
    keep = []
    acc_total = 0.0
    for it in range(n_iter):
        p = rng.normal(size=(nc, P)) / np.sqrt(minv)
        h0 = -ll + 0.5 * np.sum(p * p * minv, axis=1)
        q1, g1, p1 = q.copy(), g.copy(), p.copy()
        ll1 = ll.copy()
        for l in range(n_leapfrog):
            p1 += 0.5 * eps * g1
            q1 += eps * minv * p1
            ll1, g1 = c_oracle.batched(target, q1)
            p1 += 0.5 * eps * g1
        h1 = -ll1 + 0.5 * np.sum(p1 * p1 * minv, axis=1)
        a = np.exp(np.minimum(0.0, h0 - h1))
        a[~np.isfinite(h1)] = 0.0
        accept = rng.random(nc) < a
        q[accept], g[accept], ll[accept] = q1[accept], g1[accept], ll1[accept]
        acc_total += a.mean()
        if it >= n_warm:
            keep.append(q.copy())
    return np.array(keep), acc_total / n_iter
