import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle and the synthetic generator once (g++ only, seconds)."""
    import oracle
    from liodom_b200 import synth
    oracle.build()
    synth.build()


@pytest.fixture(scope="session")
def cuda_lib():
    """The CUDA library; GPU tests fail loudly when it is missing (no CPU fallback)."""
    from liodom_b200 import api
    return api.load()


_SEQ_CACHE = {}


def get_sequence(sensor, seed, nframes, **kw):
    from liodom_b200 import synth
    key = (sensor, seed, nframes, tuple(sorted(kw.items())))
    if key not in _SEQ_CACHE:
        _SEQ_CACHE[key] = synth.sequence(sensor, seed, nframes, **kw)
    return _SEQ_CACHE[key]


def pose_err(A, B):
    """(translation error [m], rotation angle [rad]) between two 4x4 poses."""
    dt = float(np.linalg.norm(A[:3, 3] - B[:3, 3]))
    dR = A[:3, :3] @ B[:3, :3].T
    v = np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]])
    ang = float(np.arctan2(0.5 * np.linalg.norm(v), (np.trace(dR) - 1.0) / 2.0))
    return dt, ang
