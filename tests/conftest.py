import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def rms(a):
    a = np.asarray(a, dtype=np.float64)
    return float(np.sqrt(np.mean(a * a))) if a.size else 0.0


@pytest.fixture(scope="session")
def have_ref():
    from oracle import ref
    return ref.available()


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA library, loaded; fails loudly (no CPU fallback) when it or the device is missing."""
    import dsp_b200
    dsp_b200.lib()
    if dsp_b200.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box")
    return dsp_b200
