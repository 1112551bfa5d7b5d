import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def fd5(f, x, k, h=1e-3):
    """d f / d x.flat[k] by the fourth-order five-point stencil (truncation O(h^4), round-off O(eps |f| / h)): accurate
    enough to hold analytic gradients to the 1e-8 of BASELINE.json's north star, which two-point differences are not."""
    import numpy as np
    step = h * max(1.0, abs(x.flat[k]))
    d = np.zeros(x.size)
    d[k] = step
    d = d.reshape(x.shape)
    return (-f(x + 2 * d) + 8.0 * f(x + d) - 8.0 * f(x - d) + f(x - 2 * d)) / (12.0 * step)
