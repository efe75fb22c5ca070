#!/bin/bash
# N-GPU A/B of the two-bucket gradient exchange (SVAE_SPLIT_ALLREDUCE=0/1), interleaved, C2 only
for rep in $(seq 1 ${REPS:-2}); do
  for v in 0 1; do
    echo "== SVAE_SPLIT_ALLREDUCE=$v"
    SVAE_SPLIT_ALLREDUCE=$v FLAGS="--no-extras" WD=200 bash scripts/gpu_multi.sh 2>&1 | grep -E "rc=|value|Error|error" | cut -c1-200
  done
done 2>&1 | tee gpurun_out/multi_ab.log
