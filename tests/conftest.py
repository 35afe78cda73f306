import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# TEST HOOK (lives here, not in the package): tests/test_host_drivers_mock.py re-runs GPU test files in a child pytest against the
# host-memory stand-in for the CUDA runtime (tests/cpp/cuda_mock). The child gets the mock's path in MRX_TEST_MOCK_LIB and this
# conftest points the package's loader at it before anything is loaded. The product loader itself has no such switch.
if os.environ.get("MRX_TEST_MOCK_LIB"):
    from mrcpp_b200 import _lib as _plib
    _plib.LIB_PATH = os.environ["MRX_TEST_MOCK_LIB"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _device_count():
    try:
        from mrcpp_b200 import _lib
        return int(_lib.load().mrx_device_count())
    except Exception:  # noqa: BLE001 - library not built yet: the session fixture builds it
        return -1


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a host without a GPU skips the gpu-marked tests; asking for them explicitly (`-m gpu`, what
    the GPU box runs) keeps the hard failure of the fixtures: there is no CPU fallback to fall back to."""
    asked = "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or "")
    if asked or os.environ.get("MRX_TEST_MOCK_LIB"):
        return
    if _device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (the gpu-marked tests run with -m gpu on the B200 box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def libs():
    """Build (if needed) and load the product library and the oracle."""
    from mrcpp_b200 import build
    build.build_lib()
    build.build_oracle()
    import mrcpp_b200 as mw
    from mrcpp_b200 import _lib
    _lib.init()
    import oracle_api
    return mw, oracle_api
