"""Convergence diagnostics for the gathered draws: rank-normalised split-R-hat and bulk ESS (Vehtari, Gelman, Simpson,
Carpenter, Buerkner 2021), which is what MCMCChains 6 reports through ``summarystats`` in the reference
(src/MagiJl.jl:952-961; parity unpinned there: the reference tests only check object types)."""
from __future__ import annotations

import numpy as np


def _split(x):
    """x: (n_iter, n_chains) -> (n_iter // 2, 2 * n_chains)."""
    n = x.shape[0] // 2
    return np.concatenate([x[:n], x[x.shape[0] - n:]], axis=1)


def _rank_normalise(x):
    from scipy.stats import norm, rankdata
    r = rankdata(x.reshape(-1), method="average").reshape(x.shape)
    return norm.ppf((r - 0.375) / (x.size + 0.25))


def _rhat_plain(x):
    n, m = x.shape
    cm = x.mean(axis=0)
    B = n * cm.var(ddof=1)
    W = x.var(axis=0, ddof=1).mean()
    var_plus = (n - 1) / n * W + B / n
    return float(np.sqrt(var_plus / W)) if W > 0 else float("nan")


def split_rhat(x) -> float:
    """Rank-normalised split-R-hat (max of bulk and folded) of draws x: (n_iter, n_chains)."""
    x = np.asarray(x, dtype=np.float64)
    s = _split(x)
    bulk = _rhat_plain(_rank_normalise(s))
    folded = _rhat_plain(_rank_normalise(np.abs(s - np.median(s))))
    return max(bulk, folded)


def _autocov_fft(x):
    n = x.shape[0]
    m = 1 << int(np.ceil(np.log2(2 * n)))
    xc = x - x.mean(axis=0)
    f = np.fft.rfft(xc, n=m, axis=0)
    ac = np.fft.irfft(f * np.conj(f), n=m, axis=0)[:n]
    return ac / n


def ess_bulk(x) -> float:
    """Bulk effective sample size of draws x: (n_iter, n_chains) (rank-normalised, split, Geyer initial monotone sequence)."""
    z = _rank_normalise(_split(np.asarray(x, dtype=np.float64)))
    n, m = z.shape
    if n < 4:
        return float("nan")
    acov = _autocov_fft(z)
    chain_var = acov[0] * n / (n - 1.0)
    W = chain_var.mean()
    var_plus = W * (n - 1.0) / n + (z.mean(axis=0).var(ddof=1) if m > 1 else 0.0)
    if not var_plus > 0:
        return float("nan")
    rho = 1.0 - (W - acov.mean(axis=1)) / var_plus
    rho[0] = 1.0
    # Geyer: sum of adjacent pairs, positive and monotone
    pairs = np.array([rho[2 * k] + rho[2 * k + 1] for k in range(n // 2)])
    pos = np.where(pairs < 0)[0]
    kmax = pos[0] if len(pos) else len(pairs)
    pairs = np.minimum.accumulate(pairs[:kmax]) if kmax > 0 else np.array([])
    tau = -1.0 + 2.0 * pairs.sum() if len(pairs) else 1.0
    tau = max(tau, 1.0 / np.log10(max(n * m, 10)))
    return float(n * m / tau)


def summarize(draws, names=None):
    """draws: (n_iter, n_chains, n_cols).  Returns a list of dicts (mean, sd, rhat, ess_bulk) per column."""
    d = np.asarray(draws)
    out = []
    for j in range(d.shape[2]):
        x = d[:, :, j]
        out.append(dict(name=(names[j] if names else "col%d" % j), mean=float(np.nanmean(x)), sd=float(np.nanstd(x, ddof=1)),
                        rhat=split_rhat(x) if np.all(np.isfinite(x)) else float("nan"),
                        ess_bulk=ess_bulk(x) if np.all(np.isfinite(x)) else float("nan")))
    return out
