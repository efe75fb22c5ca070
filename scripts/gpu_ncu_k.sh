#!/bin/bash
# ncu --set full of selected kernels in one eager C2 step: KREGEX=... OUT=name
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX}" -c ${COUNT:-2} -o gpurun_out/${OUT} -f \
   python bench.py --config ${CFG:-c2} --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${OUT}.log 2>&1; echo rc=$?
ncu -i gpurun_out/${OUT}.ncu-rep --page raw --csv > gpurun_out/${OUT}_raw.csv 2>/dev/null
