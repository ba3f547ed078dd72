import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The CPU suite needs the shared library for its host-only entry points (regex compiler, canonicaliser, ABI
    packer): build it in-tree when it is missing and a compiler is at hand (what __graft_entry__.build() does)."""
    from zkemail_rs_b200 import engine as _engine
    if not os.path.exists(_engine.LIB_PATH) and os.path.exists("/usr/local/cuda/bin/nvcc"):
        _engine.build_library()


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine.  GPU tests FAIL (not skip) when the library or device is missing: a silent
    fallback would void the parity claims."""
    import zkemail_rs_b200 as z
    eng = z.Engine(now_unix=1704067200)
    yield eng
    eng.close()
