import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine.  GPU tests FAIL (not skip) when the library or device is missing: a silent
    fallback would void the parity claims."""
    import zkemail_rs_b200 as z
    eng = z.Engine(now_unix=1704067200)
    yield eng
    eng.close()
