import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the oracle once per session (cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()


def have_ref():
    from oracle import ref
    return all(ref.ref_available(k) for k in ref.KINDS)


requires_ref = pytest.mark.skipif(not have_ref(), reason="oracle/_ref/*.so not built (needs /root/reference)")
