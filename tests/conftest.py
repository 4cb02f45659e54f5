import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Builds the CUDA core and the checkers once per session (no-op when up to date)."""
    import __graft_entry__ as entry
    entry.build()
    return True


@pytest.fixture(scope="session")
def gpu_ctx(built):
    from rt3_b200 import abi
    ctx = abi.Context(0)
    yield ctx
    ctx.close()


def load_golden(name):
    import numpy as np
    from rt3_b200 import abi
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    scene = None
    if "faces" in g:
        scene = abi.SceneArrays(faces=g["faces"].view(abi.FACE_DTYPE).ravel(), vertices=g["vertices"].view(abi.VERTEX_DTYPE).ravel(),
                                face_entity=g["face_entity"])
    return g, scene
