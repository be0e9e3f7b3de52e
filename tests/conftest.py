import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built_libraries():
    """Make sure the product library and the oracle are built (no-op when the .so files travelled with the snapshot)."""
    prod = os.path.join(ROOT, "hpmpc_b200", "lib", "libhpmpc_b200.so")
    if not os.path.exists(prod):
        subprocess.run(["make", "-C", os.path.join(ROOT, "hpmpc_b200", "csrc"), "-j4"], check=True, stdout=subprocess.DEVNULL)
    orc = os.path.join(ROOT, "oracle", "_ref", "liboracle.so")
    ref = os.path.join(ROOT, "oracle", "_ref", "libhpmpc_ref_c99.so")
    if not os.path.exists(orc) or (os.path.isdir("/root/reference") and not os.path.exists(ref)):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8"], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
    yield


def rel_err(a, b):
    """max over stages of |a-b| / max(1,|b|)  (the 1e-9 bar of BASELINE.json is relative; entries below 1 are
    compared absolutely so that exact zeros do not blow the ratio up)."""
    import numpy as np
    worst = 0.0
    for x, y in zip(a, b):
        x, y = np.asarray(x), np.asarray(y)
        assert x.shape == y.shape, (x.shape, y.shape)
        if y.size:
            worst = max(worst, float(np.max(np.abs(x - y) / np.maximum(1.0, np.abs(y)))))
    return worst


def rel_err_true(a, b, floor=1e-6):
    """TRUE relative error, max over the entries of |a-b| / max(|b|, floor*||b||_inf) per vector: small entries (inactive
    multipliers ~1e-8) are measured against their own size down to `floor` times the largest entry of their vector, below which
    a value is rounding noise of the sums it came from.  VERDICT r1: the max(1,|b|) metric hides 10 % errors on inactive lam."""
    import numpy as np
    worst = 0.0
    for x, y in zip(a, b):
        x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
        assert x.shape == y.shape, (x.shape, y.shape)
        if y.size:
            den = np.maximum(np.abs(y), floor * float(np.max(np.abs(y))) + 1e-300)
            worst = max(worst, float(np.max(np.abs(x - y) / den)))
    return worst
