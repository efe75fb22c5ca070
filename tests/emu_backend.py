"""pytest plumbing for tests/simt_emu: load the host build of the SIMT kernel sources behind the product's own
Python layer, so that spatial_vae.functional / models / trainer can be driven with CPU tensors IN TESTS.

Test infrastructure only: the package itself never loads this library (spatial_vae._lib opens libsvae_b200.so and
nothing else) and keeps refusing CPU tensors; the patches below exist for the duration of one test.
"""
import ctypes as C

import torch

from tests.simt_emu.build import build


def load_emu(fresh_copy_dir=None):
    """fresh_copy_dir: load a private copy of the library (its function-local statics, e.g. the environment
    switches read once per process, start fresh)."""
    import spatial_vae._lib as L
    path = build()
    if fresh_copy_dir is not None:
        import shutil
        path = shutil.copy(path, str(fresh_copy_dir))
    return L.declare(C.CDLL(path))


def install(monkeypatch, fresh_copy_dir=None):
    """Route the ctypes calls of spatial_vae.functional to the host build and let CPU tensors through."""
    import spatial_vae._lib as L
    import spatial_vae.functional as SF
    emu = load_emu(fresh_copy_dir)
    monkeypatch.setattr(L, "lib", emu)
    monkeypatch.setattr(SF, "_require_cuda", lambda *t: None)
    monkeypatch.setattr(SF, "_stream", lambda: 0)
    buffers = {}

    def workspace(nbytes, device):
        ws = buffers.get("ws")
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(int(nbytes) + 4096, dtype=torch.uint8)
            buffers["ws"] = ws
        ws.fill_(0xFF)        # NaN in fp32 and bf16: a kernel that reads scratch it never wrote poisons the result
        return ws

    monkeypatch.setattr(SF, "workspace", workspace)
    return emu


def install_all(monkeypatch, fresh_copy_dir=None):
    """install() plus stand-ins for the CUDA-only HOST machinery the drivers use (device selection, pinned memory,
    events, CUDA graphs), so that whole command lines can run on the emulation.  Returns the CPU device."""
    install(monkeypatch, fresh_copy_dir)
    cpu = torch.device("cpu")
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
    return cpu
