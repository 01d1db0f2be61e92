import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCENES = os.path.join(ROOT, "scenes")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def scene_path(name):
    return os.path.join(SCENES, name + ".gltf")


@pytest.fixture(scope="session")
def rt():
    import rtb200
    if not os.path.exists(rtb200.LIB_PATH):
        rtb200.build()
    return rtb200


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def gpu_rt(rt):
    """The product library with a CUDA device present.  GPU tests fail (not skip) if the device is missing."""
    n = rt.device_count()
    assert n > 0, "no CUDA device visible: -m gpu tests must run on the GPU box"
    return rt
