"""The C restatement (CPU baseline, oracle/magi_oracle.c) must agree with the numpy oracle."""
import numpy as np
import pytest

from oracle import c_oracle, magi_oracle as mo
from tests import helpers as H


@pytest.mark.parametrize("model,n,b,kw", [("fn", 41, 6, {}), ("fn", 201, 20, {"beta": (1.0, 2.0, 5.0)}), ("hes1", 33, 5, {}), ("lv", 50, 49, {}),
                                          ("fn", 30, 4, {"sigma_fixed": True}), ("fn", 1, 0, {"obs_every": 1})])
def test_c_port_matches_numpy_oracle(model, n, b, kw):
    prob = H.make_problem(model=model, n=n, b=b, n_chains=4, seed=n, **kw)
    ll, g = c_oracle.batched(prob["target"], prob["params"], nthreads=2)
    ll_ref, g_ref = H.oracle_batched(prob)
    H.assert_parity(ll, g, ll_ref, g_ref, "C port %s n=%d" % (model, n))


def test_c_port_guards():
    prob = H.make_problem(n=21, b=4, n_chains=3, seed=2)
    p = prob["params"].copy()
    p[1, 3] = np.nan
    ll, g = c_oracle.batched(prob["target"], p)
    assert ll[1] == -np.inf and not g[1].any() and np.isfinite(ll[0]) and np.isfinite(ll[2])
