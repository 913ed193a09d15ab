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


def _have_gpu():
    try:
        from fredholm_b200.api import lib
        return lib().fr_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_mod():
    """The host oracle (reference sources compiled for the CPU).  Test infrastructure."""
    from oracle import binding
    if not binding.available():
        binding.build()
    if not binding.available():
        pytest.skip("oracle library not built (needs /root/reference)")
    return binding


@pytest.fixture()
def oracle(oracle_mod):
    return oracle_mod.Oracle()


@pytest.fixture()
def renderer():
    from fredholm_b200 import Renderer
    r = Renderer(0)
    yield r
    r.close()


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def rel_mse(x, ref):
    """mean over pixels and channels of (x-ref)^2 / (ref^2 + 1e-2)."""
    x = np.asarray(x, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.mean((x - ref) ** 2 / (ref ** 2 + 1e-2)))
