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
    import torch
    from tests import emu_backend
    emu_backend.install(monkeypatch)
    cpu = torch.device("cpu")
    for helper in ("_cuda", "_dev"):
        if hasattr(request.module, helper):
            monkeypatch.setattr(request.module, helper, lambda: cpu)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)

    class _Event:
        def __init__(self, *a, **k): pass
        def record(self, *a, **k): pass
        def synchronize(self): pass
        def elapsed_time(self, other): return 0.0

    monkeypatch.setattr(torch.cuda, "Event", _Event)
    import spatial_vae.driver as D
    monkeypatch.setattr(D, "pick_device", lambda *a, **k: cpu)
    from spatial_vae.trainer import Trainer
    monkeypatch.setattr(Trainer, "step_graphed", lambda self, *a, **k: self.step(*a, **k))   # no CUDA graphs on a host
    yield
