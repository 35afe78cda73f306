import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


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
