#!/bin/bash
# A/B on the same box: $1 = env assignment for variant B
mkdir -p gpurun_out
for i in 1 2; do
  python bench.py --steps 200 --warmup 5 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('A', round(d['ms_per_step'],3), d['roofline']['all_kernels_ms'])"
  env $1 python bench.py --steps 200 --warmup 5 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B', round(d['ms_per_step'],3), d['roofline']['all_kernels_ms'])"
done
