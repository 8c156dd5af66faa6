import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_model(name):
    """-> (state_dict, cfg, arrays) for the whole-model fixtures."""
    g = load_golden(name)
    sd = {k[4:]: v for k, v in g.items() if k.startswith("sd::")}
    cfg = json.loads(bytes(g["cfg_json"]).decode())
    return sd, cfg, g


@pytest.fixture(scope="session")
def golden():
    return load_golden


def assert_eig_close(actual, desired, rtol=1e-5, what="eigenvalues"):
    """The north-star tolerance (rel 1e-5 in fp32) applied where it is well-posed.  lambda = exp(dt*A) has
    condition number |ln lambda| with respect to its exponent, so an fp32-exact exponent (rel 1e-7.. 1e-6)
    gives  |d lambda| / lambda ~ |ln lambda| * eps  -- the reference's own fp32 path differs from fp64 truth by
    3.6e-5 at lambda = 2e-5.  The bound used everywhere: |a - d| <= rtol * |d| * (1 + |ln |d||)."""
    a = np.asarray(actual, np.float64)
    d = np.asarray(desired, np.float64)
    assert a.shape == d.shape, (a.shape, d.shape)
    fin = np.isfinite(d)
    assert (np.isfinite(a) == fin).all(), what + ": finite masks differ"
    with np.errstate(divide="ignore"):
        cond = 1.0 + np.abs(np.log(np.maximum(np.abs(d[fin]), 1e-300)))
    err = np.abs(a[fin] - d[fin])
    bound = rtol * np.abs(d[fin]) * cond + 1e-44
    bad = err > bound
    assert not bad.any(), "%s: %d/%d outside rel %g (worst ratio %.3g)" % (what, bad.sum(), bad.size, rtol, (err / bound).max())
