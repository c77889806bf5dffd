import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def built():
    """Build the oracle (always) and the CUDA library (cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    ge.build()
    return True


def bits(a):
    import numpy as np
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_bit_equal(a, b, what=""):
    import numpy as np
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    if a.size == 0:
        return
    eq = bits(a) == bits(b)
    if a.dtype == np.float32:
        eq = eq | (np.isnan(a) & np.isnan(b))
    assert bool(eq.all()), "%s: %d of %d elements differ (first at %s)" % (what, int((~eq).sum()), a.size, np.argwhere(~eq)[0])
