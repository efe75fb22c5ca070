#!/bin/bash
# Round 2, first GPU call: validated suite, the chunk / dual-stream experiments of scripts/gpu_round2_first.sh,
# a launch list of the C2 step and ncu --set full captures of the three bandwidth-bound SIMT passes.
mkdir -p gpurun_out
{
  echo "== gpu suite"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  echo "== C2 default"; timeout 300 python bench.py --config c2 --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/r2c1_c2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', d['roofline']['all_kernels_ms'])"
  echo "== chunk sweep"
  for ch in 256 96 48 24; do
    timeout 300 python bench.py --config c2 --chunk $ch --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk=$ch', round(d['ms_per_step'],3), 'ms/step')"
  done
  echo "== dual stream"
  for ch in 0 96; do
  SVAE_DUAL_STREAM=1 timeout 300 python bench.py --config c2 --chunk $ch --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dual chunk=$ch', round(d['ms_per_step'],3), 'ms/step')"
  done
  echo "== launch list (C2, eager, 2 steps)"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_base_launches.csv \
     python bench.py --config c2 --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1; echo rc=$?
  echo "== ncu --set full: layer0 / out_backward / image_col_reduce"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"layer0_k|out_backward|image_col_reduce" -c 3 \
     -o gpurun_out/r02_simt_full python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1; echo rc=$?
} 2>&1 | tee gpurun_out/r2_call1.log
