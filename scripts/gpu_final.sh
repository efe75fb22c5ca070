#!/bin/bash
# what the driver runs at round end, in one call: the -m gpu suite, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
{
  echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
  echo "== bench"; timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "rc=$?"
  python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/final_bench.json") if l.startswith("{")][-1])
print({k: d.get(k) for k in ("metric", "value", "unit", "ms_per_step", "steps", "gpu_launches", "vs_baseline")})
print("e2e", d["e2e"]); print("roofline", d["roofline"]); print("cpu_baseline", d["cpu_baseline"]); print("clocks", d["clocks"])
for k, v in d.get("extra_configs", {}).items(): print(k, v)
print("precision_modes", d.get("precision_modes")); print("gpu_eager_baseline", d.get("gpu_eager_baseline"))
PY
  echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/final_ref.err | tee gpurun_out/final_ref.json | cut -c1-400; echo "rc=$?"
} 2>&1 | tee gpurun_out/final.log
