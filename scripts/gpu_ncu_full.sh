#!/bin/bash
# ncu --set full on the tcgen05 GEMMs (after the same command exited 0 without ncu)
mkdir -p gpurun_out
python scripts/gemm_only.py > gpurun_out/plain_gemm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 3 -f -o gpurun_out/gemm_full python scripts/gemm_only.py > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
