#!/bin/bash
# same-box sweep of environment assignments (one per line in $1), baseline interleaved: bench lines for $CFGS
mkdir -p gpurun_out
while read -r assign; do
  echo "== ${assign:-baseline}" | tee -a gpurun_out/ab.log
  env $assign bash scripts/gpu_ab.sh
done < "$1"
