import glob
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


def golden_names(prefix=""):
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


_cache = {}


def load_golden(name):
    """-> (scene, size, linear_rgb, npz)"""
    if name not in _cache:
        from svgrasterize_b200 import sceneio

        z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        _cache[name] = (sceneio.load_scene(z), tuple(float(v) for v in z["size"]), bool(z["linear_rgb"]), z)
    return _cache[name]


@pytest.fixture(scope="session")
def golden():
    return load_golden
