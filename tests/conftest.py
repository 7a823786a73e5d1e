import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure): built on demand with gcc."""
    from oracle import oracle as O
    O.build()
    O.lib()
    return O


@pytest.fixture(scope="session")
def vpl():
    """The product package (ctypes over libvplines_b200.so)."""
    return importlib.import_module("vplines_slam_b200")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("vplines-slam_b200.synth")


@pytest.fixture(scope="session")
def mh04():
    return np.load(os.path.join(GOLDEN, "mh04_frames.npz"))["frames"]


@pytest.fixture(scope="session")
def golden():
    return {n: np.load(os.path.join(GOLDEN, n + ".npz")) for n in ("cv2_lsd", "cv2_prims", "cv2_hamming")}


def have_gpu():
    try:
        m = importlib.import_module("vplines_slam_b200")
        return m.capi.device_count() > 0
    except Exception:
        return False
