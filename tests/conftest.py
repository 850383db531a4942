import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

# north_star tolerances: |y - y_ref|_i <= tol * sum_j |a_ij x_j|
TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def assert_within_bound(y, y_ref, scale, dtype, what=""):
    """The north_star criterion, element-wise.  scale_i = sum_j |a_ij x_j| (oracle.abs_rowsum)."""
    tol = TOL[np.dtype(dtype)]
    y = np.asarray(y, dtype=np.float64)
    y_ref = np.asarray(y_ref, dtype=np.float64)
    assert y.shape == y_ref.shape, what
    assert np.all(np.isfinite(y)), what + ": non-finite result"
    err = np.abs(y - y_ref)
    bound = tol * np.asarray(scale, dtype=np.float64)
    bad = np.nonzero(err > bound)[0]
    assert bad.size == 0, "%s: %d rows outside %g*sum|a_ij x_j| (first row %d: err %g, bound %g)" % (
        what, bad.size, tol, bad[0], err[bad[0]], bound[bad[0]])


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def hs():
    import heterogeneous_spmv_b200 as m
    m.lib()
    return m


@pytest.fixture(scope="session")
def gen():
    from heterogeneous_spmv_b200 import generators
    return generators
