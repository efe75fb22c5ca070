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
    monkeypatch.setenv("SVAE_UNVALIDATED_OPTIONS", "1")
    return emu
