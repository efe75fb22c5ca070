import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "spatial-vae_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


# ---- SVAE_TEST_BACKEND=emu: run the bodies of the `gpu` tests on the CPU against tests/simt_emu -------------------
# (the library's SIMT kernel sources compiled for the host; see tests/simt_emu/cuda_emu.h).  Used by
# tests/test_emu_gpu_suite.py to execute, before any GPU time is spent, the very tests the B200 box will run.
# Not a product mode: the package itself never loads the host build.
EMU_BACKEND = os.environ.get("SVAE_TEST_BACKEND") == "emu"


@pytest.fixture(autouse=True)
def _emu_backend(request, monkeypatch):
    if not EMU_BACKEND or request.node.get_closest_marker("gpu") is None:
        yield
        return
    from tests import emu_backend
    cpu = emu_backend.install_all(monkeypatch)
    for helper in ("_cuda", "_dev"):
        if hasattr(request.module, helper):
            monkeypatch.setattr(request.module, helper, lambda: cpu)
    yield
