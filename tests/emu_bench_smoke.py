"""Manual check (about 1.5 minutes, not part of the pytest suites): bench.py's measured arm, start to finish, on the
emulation of tests/simt_emu -- the epoch permutation, launch counting, the timed loop, the end-to-end loop, the GEMM
timing section and the JSON line.  The numbers mean nothing (events are faked, the "device" is the CPU); the point is
that every line of bench.py's main path executes and the JSON carries every key of the contract.
    python tests/emu_bench_smoke.py
"""
import contextlib
import io
import json
import sys, os, types
os.environ["BENCH_NO_SAMPLER"]="1"
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/spatial-vae_b200')
import pytest, torch
from tests import emu_backend
mp = pytest.MonkeyPatch(); emu_backend.install_all(mp)
real_device = torch.device
class FakeDevice:
    def __new__(cls, *a, **k): return real_device("cpu")
torch.device = FakeDevice
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a, **k: None
class _Ev:
    def __init__(self,*a,**k): pass
    def record(self,*a,**k): pass
    def synchronize(self): pass
    def elapsed_time(self, o): return 5.0
torch.cuda.Event = _Ev
torch.cuda.current_stream = lambda *a, **k: types.SimpleNamespace(cuda_stream=0)
real_gen = torch.Generator
torch.Generator = lambda device=None: real_gen()
real_randperm = torch.randperm
import bench
sys.argv = ["bench.py", "--config", os.environ.get("BENCH_SMOKE_CONFIG", "c1"), "--batch", "1", "--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--no-extras"]
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
line = json.loads(buf.getvalue().strip().splitlines()[-1])
for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
    assert key in line, key
assert "workload" in line["config"] and line["gpu_launches"] > 0
assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
print("bench.py main path ok on the emulation:", {k: line[k] for k in ("metric", "unit", "gpu_launches", "dtype")})
