import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref():
    """The compiled UNMODIFIED reference + harness (oracle/_ref/libgnxref.so)."""
    import _harness
    if not os.path.exists(_harness.REF_LIB):
        pytest.skip("oracle/_ref/libgnxref.so not built (needs /root/reference at build time)")
    return _harness.Ref()


@pytest.fixture(scope="session")
def emul():
    """The product's per-path device functions compiled for the host (tests/emul)."""
    import _harness
    from gnxraytracer_b200.build import build_emul
    build_emul()
    return _harness.Emul()
