import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
RIG17 = os.path.join(GOLDEN, "rig17.json")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def rig_path():
    return RIG17


@pytest.fixture(scope="session")
def ctx():
    from surround360_b200 import api
    c = api.Context(0)
    c.load_rig(RIG17)
    yield c
    c.close()


@pytest.fixture(scope="session")
def orig():
    import oracle_bind as ob
    return ob.Rig.load(RIG17)


@pytest.fixture(scope="session")
def frames17():
    """17 synthetic 2048^2 frames (cached on disk between test sessions)."""
    import numpy as np
    from surround360_b200 import synth
    cache = "/tmp/s360_frames17_v1.npy"
    if os.path.exists(cache):
        return list(np.load(cache))
    fr = synth.render_rig_frames(RIG17)
    try:
        np.save(cache, np.stack(fr))
    except OSError:
        pass
    return fr
