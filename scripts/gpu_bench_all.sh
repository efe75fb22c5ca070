#!/bin/bash
mkdir -p gpurun_out
for c in c1 c3 c4 c5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.log 2>&1; echo "$c exit $?"; tail -c 1500 gpurun_out/bench_$c.log | tail -2
done
