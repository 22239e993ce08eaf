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


def _load(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def dtm188():
    return _load("dtm188.npz")


@pytest.fixture(scope="session")
def small_cases():
    z = _load("small_cases.npz")
    n = int(z["ncases"])
    cases = []
    for i in range(n):
        pre = "c%02d_" % i
        cases.append({k[len(pre):]: v for k, v in z.items() if k.startswith(pre)})
    flows = []
    for i in range(int(z["nflow"])):
        pre = "f%02d_" % i
        flows.append({k[len(pre):]: v for k, v in z.items() if k.startswith(pre)})
    return cases, flows


@pytest.fixture(scope="session")
def fractal256():
    return _load("fractal256.npz")
