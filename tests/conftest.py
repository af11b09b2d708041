import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


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
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    class G:
        def __init__(self):
            self._c = {}

        def __call__(self, name):
            if name not in self._c:
                self._c[name] = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
            return self._c[name]

    return G()


@pytest.fixture(scope="session", autouse=True)
def _native_library_is_built():
    """The test session (not the product) compiles libbfa_b200.so when it is missing or stale and nvcc is at hand -- the same
    thing __graft_entry__.build() does; the product itself never builds or falls back: _cabi.lib() raises without the .so."""
    import importlib.util
    import shutil
    root = Path(__file__).resolve().parents[1]
    spec = importlib.util.spec_from_file_location("bfa_build_for_tests", root / "bournemouth-forced-aligner_b200" / "build.py")
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    if b._stale() and shutil.which("nvcc"):
        b.build()
    yield


@pytest.fixture(scope="session")
def dev():
    import torch
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def bfa():
    import bfa_b200
    return bfa_b200


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    return oracle
