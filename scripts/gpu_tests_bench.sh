#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x > gpurun_out/t_all.log 2>&1
echo "tests exit $?" >> gpurun_out/t_all.log
tail -n 25 gpurun_out/t_all.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_c2.log; tail -3 gpurun_out/bench_c2.log
