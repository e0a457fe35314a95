"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/magi_b200.h
declares, and refuses to compute without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "magi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(magi_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(pkg):
    from manifold_constrained_gaussian_process_inference_b200 import _lib
    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 14
    for s in declared:
        assert hasattr(L, s), "libmagi_b200.so does not export %s" % s
    assert set(_lib.EXPORTED) <= set(declared)
    assert L.magi_version() >= 100


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import, link or execute it."""
    pk = os.path.join(ROOT, "manifold_constrained_gaussian_process_inference_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|magi_oracle|libmagi_oracle)", txt), f


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from manifold_constrained_gaussian_process_inference_b200 import _lib
    with pytest.raises(_lib.MagiError):
        pkg.MagiTarget.from_config(np.zeros((3, 2)), np.arange(3.0), np.ones((2, 2)), pkg.fn_system(), [0.1, 0.1])
    with pytest.raises(_lib.MagiError):
        pkg.calculate_gp_covariances(pkg.GPCov(), pkg.create_rbf_kernel(1.0, 1.0), [1.0, 1.0], np.arange(3.0), 1, complexity=2)


def test_host_side_mirrors(pkg):
    import numpy as np
    k = pkg.create_matern52_kernel(2.0, 1.5)
    assert (k.kind, k.variance, k.lengthscale) == ("matern52", 2.0, 1.5)
    with pytest.raises(AssertionError):
        pkg.create_rbf_kernel(-1.0, 1.0)
    s = pkg.fn_system()
    assert (s.n_dims, s.thetaSize, s.model_id) == (2, 3, 0)
    A = np.arange(16.0).reshape(4, 4)
    B = pkg.mat2band(A, 1, 0)
    assert B[1, 0] == A[1, 0] and B[0, 1] == 0 and B[2, 0] == 0
    assert pkg.capabilities(None) == pkg.LogDensityOrder(1)
