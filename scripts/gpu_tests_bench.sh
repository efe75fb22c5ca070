#!/bin/bash
# GPU pass: tcgen05 GEMM unit tests first (short timeout), then everything, then the bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=120 -x -k "tc_gemm" > gpurun_out/t_gemm.log 2>&1
echo "gemm tests exit $?" >> gpurun_out/t_gemm.log
tail -n 12 gpurun_out/t_gemm.log
if grep -q "gemm tests exit 0" gpurun_out/t_gemm.log; then
  timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x > gpurun_out/t_all.log 2>&1
  echo "tests exit $?" >> gpurun_out/t_all.log
  tail -n 12 gpurun_out/t_all.log
  timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_c2.log; tail -3 gpurun_out/bench_c2.log
  SVAE_TC_CTA_GROUP=1 timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_cg1.log 2>&1; tail -1 gpurun_out/bench_c2_cg1.log
fi
