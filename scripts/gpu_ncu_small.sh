#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sgemm_kernel -s 12 -c 6 -f -o gpurun_out/sgemm_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_small.log 2>&1
echo "exit $?"
