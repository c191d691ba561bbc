import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    """The product shared library; (re)built in-tree if missing (nvcc needs no GPU)."""
    from partgs_b200 import _lib, build
    if not _lib.LIB_PATH.exists():
        build.build()
    return _lib.load()


def pytest_runtest_logreport(report):
    """PGS_INSTAFAIL=1: print a failure's traceback as soon as the test ends (short GPU calls under a hard time limit
    keep what was printed, not what pytest would have summarised at the end)."""
    import os
    if os.environ.get("PGS_INSTAFAIL") and report.failed:
        print("\n[instafail] " + report.nodeid + "\n" + str(report.longrepr)[-3000:], flush=True)
