import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import lvo
    return lvo.Oracle("own")


@pytest.fixture(scope="session")
def oracle_ref():
    """The oracle drivers on the reference's madmann91/bvh library (oracle/_ref, prebuilt from /root/reference)."""
    from oracle import lvo
    try:
        return lvo.Oracle("ref")
    except (FileNotFoundError, OSError):
        pytest.skip("oracle/_ref not built (needs /root/reference/submodules/bvh)")


@pytest.fixture(scope="session")
def ctx():
    import linevis_b200 as lv
    c = lv.Context(0)
    yield c
    c.close()
