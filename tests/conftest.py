import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_rms(a, ref):
    """sqrt(sum |a-ref|^2 / sum |ref|^2), SURVEY 8(d)."""
    import numpy as np
    a = np.asarray(a, np.float64); ref = np.asarray(ref, np.float64)
    return float(np.sqrt(np.sum((a - ref) ** 2) / np.sum(ref ** 2)))


@pytest.fixture(scope="session")
def engine():
    from starflood_b200 import engine as e
    e.init()
    yield e
    e.shutdown()
