#!/bin/bash
# round-2 profile pass: launch list of one eager C2 step + ncu --set full of the four big kernels inside that step
mkdir -p gpurun_out
TAG=${TAG:-r02_v3}
python bench.py --config c2 --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-extras > /dev/null 2>&1 || echo "plain run failed"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --config c2 --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-extras > gpurun_out/${TAG}_launches.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"dw_xf_kernel|dx_red_kernel|layer0_k|tc_gemm_kernel<\(int\)0, \(int\)0, \(int\)1" -s 16 -c 4 \
   -o gpurun_out/${TAG}_full -f python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-extras > gpurun_out/${TAG}_full.log 2>&1; echo rc=$?
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
