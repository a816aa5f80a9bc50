import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # checkers: CPU oracle (always) and, where the reference tree is mounted, its own kernels
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    # product library: make is a no-op when it is up to date (nvcc is needed only when sources changed)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cudacam_b200", "csrc")])


def _have_gpu():
    try:
        from cudacam_b200 import _lib
        return _lib.lib.b2c_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def have_gpu():
    return _have_gpu()


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here; runs on the B200 box")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
