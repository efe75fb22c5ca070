"""In-situ kernel timeline of the C2 train step (eager launches, torch.profiler / CUPTI): per-kernel device time inside a
running loop (sustained clocks), launch gaps included.  python scripts/step_profile.py [config] [graph]"""
import os, sys, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spatial-vae_b200")]
import torch
import bench
from torch.profiler import profile, ProfilerActivity

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
use_graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
c = dict(bench.CONFIGS[name])
dev = torch.device("cuda", 0)
precision = os.environ.get("SVAE_PROFILE_PRECISION", "fast")     # fast | parity_tc | parity
w = bench.Workload(name, c, dev, 0, 1, precision)
fn = w.trainer.step_graphed if use_graph else w.trainer.step
for _ in range(30):
    w.device_step(fn)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(20):
        w.device_step(fn)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
for e in ev:
    k = e.name[:60]
    a = agg.setdefault(k, [0.0, 0]); a[0] += e.time_range.elapsed_us(); a[1] += 1
span = ev[-1].time_range.end - ev[0].time_range.start
busy = sum(v[0] for v in agg.values())
print(f"20 steps: span {span/20:.1f} us/step, kernels busy {busy/20:.1f} us/step (sum of durations; overlap counts twice)")
for k, (t, n) in sorted(agg.items(), key=lambda x: -x[1][0])[:32]:
    print(f"{t/20:9.1f} us/step x{n/20:5.1f}  {k}")

# one step's timeline (start offset, duration, stream) when asked: python scripts/step_profile.py c2 graph timeline
if len(sys.argv) > 3 and sys.argv[3] == "timeline":
    starts = [i for i, e in enumerate(ev) if "gather_rows_k" in e.name]
    a, b = starts[-2], starts[-1]
    t0 = ev[a].time_range.start
    print("\none step: start us | duration us | stream | kernel")
    for e in ev[a:b]:
        stream = getattr(e, "device_resource_id", None)
        print(f"{e.time_range.start - t0:9.1f} {e.time_range.elapsed_us():8.1f}  s{stream}  {e.name[:90]}")
