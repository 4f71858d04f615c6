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


@pytest.fixture(scope="session")
def port():
    """The C restatement oracle (oracle/ngt_oracle.c)."""
    from oracle import pyoracle
    return pyoracle.Port()


@pytest.fixture(scope="session")
def sift5k():
    return dict(np.load(os.path.join(GOLDEN, "sift5k.npz")))


@pytest.fixture(scope="session")
def synth_golden():
    z = dict(np.load(os.path.join(GOLDEN, "synth.npz")))
    for k in [k for k in z if k.endswith("_objects_alias")]:
        z[k[: -len("_alias")]] = z[str(z[k]) + "_objects"]
    return z


@pytest.fixture(scope="session")
def eng():
    """The device engine over the C ABI (GPU tests only)."""
    from ngt_b200 import engine
    return engine
