#!/bin/bash
# per-kernel times of one eager step: CFG=c2 OUT=name
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${OUT}.csv \
   python bench.py --config ${CFG:-c2} --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${OUT}.log 2>&1; echo rc=$?
