import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.ksoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The compiled, unmodified reference (oracle/_ref).  Built here from /root/reference;
    on the GPU box the prebuilt .so travels with the snapshot."""
    from oracle.ksoracle import Ref
    try:
        return Ref()
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (reference source absent)")
