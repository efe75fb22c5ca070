#!/bin/bash
# One call: launch list of a short bench run + ncu --set full of the three tcgen05 GEMMs (each after the same
# command exited 0 without ncu).
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "launch list exit $?"
python scripts/gemm_only.py > gpurun_out/plain_gemm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 3 -f -o gpurun_out/gemm_full python scripts/gemm_only.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
