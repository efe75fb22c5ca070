#!/bin/bash
# same-box A/B of two builds of the library: spatial_vae/alt_old.so against spatial_vae/alt_new.so, interleaved
L=spatial-vae_b200/spatial_vae
for rep in 1 2; do
  for v in old new; do
    cp $L/alt_$v.so $L/libsvae_b200.so
    echo "== $v" | tee -a gpurun_out/ab.log
    bash scripts/gpu_ab.sh
  done
done
cp $L/alt_new.so $L/libsvae_b200.so
