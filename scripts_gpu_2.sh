#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/t_all.log 2>&1
echo "tests exit $?" >> gpurun_out/t_all.log
tail -n 15 gpurun_out/t_all.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_c2.log; tail -5 gpurun_out/bench_c2.log
timeout 900 python bench.py --steps 3 --warmup 3 --precision parity --no-cpu-baseline --batch 128 > gpurun_out/bench_c2_parity.log 2>&1; tail -3 gpurun_out/bench_c2_parity.log
