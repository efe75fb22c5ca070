#!/bin/bash
# A/B timing helper: bench lines (ms/step) for the configs in $CFGS, no CPU baseline, no extras
mkdir -p gpurun_out
for cfg in ${CFGS:-c2}; do
  timeout 300 python bench.py --config $cfg --steps ${STEPS:-100} --warmup 5 --no-cpu-baseline --no-extras 2>gpurun_out/bench_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', round(d['ms_per_step'],4), 'ms/step', round(d['value']), d['last_step'])" || tail -3 gpurun_out/bench_err.log
done 2>&1 | tee -a gpurun_out/ab.log
