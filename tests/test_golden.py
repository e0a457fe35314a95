"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle must keep reproducing them on
CPU, and the CUDA path must match them on the GPU without consulting the oracle at run time."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import magi_oracle as mo
from tests import helpers as H

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
NAMES = {0: "fn", 1: "hes1", 7: "lv"}


def _load(path):
    z = np.load(path)
    b = int(z["bandsize"])
    covs = []
    for d in range(z["bands"].shape[0]):
        g = mo.GPCov(bandsize=b, tvec=z["tvec"])
        g.CinvBand, g.mphiBand, g.KinvBand = z["bands"][d]
        covs.append(g)
    return z, covs


def test_fixtures_present():
    assert len(GOLD) >= 7
    k = json.load(open(os.path.join(os.path.dirname(GOLD[0]), "reference_known_answers.json")))
    assert k["missing_observation_gradient_delta"]["value"] == 1.0


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_reproduces_golden(path):
    z, covs = _load(path)
    tgt = mo.make_target(z["yobs"], covs, int(z["model_id"]), z["sigma_init"], tuple(z["beta"]), bool(z["sigma_is_fixed"]))
    for c in range(min(2, z["params"].shape[0])):
        ll, g = mo.logdensity_and_gradient(tgt, z["params"][c])
        H.assert_parity([ll], g, [z["ll"][c]], z["grad"][c], os.path.basename(path))


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_cuda_matches_golden(pkg, path):
    z, covs = _load(path)
    pc = [pkg.GPCov(tvec=z["tvec"], bandsize=int(z["bandsize"]), CinvBand=c.CinvBand, mphiBand=c.mphiBand, KinvBand=c.KinvBand) for c in covs]
    n, D = z["yobs"].shape
    sysm = pkg.get_ode_system(NAMES[int(z["model_id"])])
    tg = pkg.MagiTarget(z["yobs"], pc, sysm, z["sigma_init"], list(z["beta"]), n, D, sysm.thetaSize, bool(z["sigma_is_fixed"]))
    ll, g = tg.logdensity_and_gradient_batched(z["params"])
    H.assert_parity(ll, g, z["ll"], z["grad"], os.path.basename(path))
