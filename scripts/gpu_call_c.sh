#!/bin/bash
mkdir -p gpurun_out
{
  echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  for v in 0 1; do
    SVAE_SIDE_STREAM=$v timeout 300 python bench.py --config c2 --steps 200 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2 side=$v', round(d['ms_per_step'],4), 'ms/step', round(d['value']), 'e2e', round(d['e2e']['value']))"
    SVAE_SIDE_STREAM=$v timeout 300 python bench.py --config c2 --steps 200 --warmup 5 --no-cpu-baseline --no-extras --no-graph 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2 eager side=$v', round(d['ms_per_step'],4), 'ms/step')"
    SVAE_SIDE_STREAM=$v timeout 300 python bench.py --config c1 --steps 200 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c1 side=$v', round(d['ms_per_step'],4), 'ms/step', round(d['value']))"
  done
} 2>&1 | tee gpurun_out/call_c.log
