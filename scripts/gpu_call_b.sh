#!/bin/bash
mkdir -p gpurun_out
{
  echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
  echo "== CTF fast kernel parity"; SVAE_CTF_FAST=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_options.py -m gpu -q -k "ctf" 2>&1 | tail -3
  for v in 0 1; do
    SVAE_CTF_FAST=$v timeout 300 python bench.py --config c5 --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 ctf_fast=$v', round(d['ms_per_step'],3), 'ms/step', round(d['value']))"
  done
  timeout 300 python bench.py --config c2 --steps 100 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2', round(d['ms_per_step'],3), 'ms/step', round(d['value']), d['launches_per_step'])"
} 2>&1 | tee gpurun_out/call_b.log
