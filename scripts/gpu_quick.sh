#!/bin/bash
# quick GPU check: new-kernel unit tests, the gpu suite, C2 bench (no CPU baseline)
mkdir -p gpurun_out
{
  echo "== unit"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${UNIT:-tc_}" 2>&1 | tail -4
  echo "== gpu suite"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
  for cfg in ${CFGS:-c2}; do
  echo "== bench $cfg"; timeout 300 python bench.py --config $cfg --steps 100 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_err.log | tee gpurun_out/quick_$cfg.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', round(d['value']), d.get('roofline',{}).get('all_kernels_ms'), d['last_step'])"
  done
  tail -3 gpurun_out/bench_err.log
} 2>&1 | tee gpurun_out/quick.log
