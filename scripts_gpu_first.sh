#!/bin/bash
# first GPU pass: SIMT (parity precision) tests, then the tcgen05 GEMM unit tests, then the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 -k "parity_precision or chunking or empty or inference or (c1_shape and parity) or (galaxy and parity) or (ctf_40 and parity) or (module_forward and parity) or (adam and parity)" > gpurun_out/t1_parity.log 2>&1
echo "t1 exit $?" >> gpurun_out/t1_parity.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 -k "tc_gemm_forward" > gpurun_out/t2_fwd.log 2>&1
echo "t2 exit $?" >> gpurun_out/t2_fwd.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 -k "tc_gemm_dx" > gpurun_out/t3_dx.log 2>&1
echo "t3 exit $?" >> gpurun_out/t3_dx.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 -k "tc_gemm_dw" > gpurun_out/t4_dw.log 2>&1
echo "t4 exit $?" >> gpurun_out/t4_dw.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 -k "fast" > gpurun_out/t5_fast.log 2>&1
echo "t5 exit $?" >> gpurun_out/t5_fast.log
tail -n 30 gpurun_out/t1_parity.log gpurun_out/t2_fwd.log gpurun_out/t3_dx.log gpurun_out/t4_dw.log gpurun_out/t5_fast.log
