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
    """The context the -m gpu tests run on.  LV_EMU_CTX=1 swaps in the host emulation of the library (tests/emu): the GPU parity tests
    themselves can then be rehearsed without a GPU (`LV_EMU_CTX=1 pytest -m gpu tests/test_gpu_parity.py`; slow, skip the full-size file)."""
    import linevis_b200 as lv
    lib_path = None
    if os.environ.get("LV_EMU_CTX"):
        import importlib.util
        spec = importlib.util.spec_from_file_location("build_emu", os.path.join(ROOT, "tests", "emu", "build_emu.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        lib_path = mod.build()
    c = lv.Context(0, lib_path=lib_path)
    yield c
    c.close()
