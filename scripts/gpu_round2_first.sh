#!/bin/bash
# First GPU call of round 2 (one GPU): everything that was written after round 1's GPU budget ran out.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_round2_first.sh'
# 1. the validated suite must still be green (two null-guarded branches were added to validated kernels);
# 2. the option kernels + softplus / activation tests that have only run under tests/simt_emu so far;
# 3. the register-tiled 39x39 CTF kernel: parity, then A/B on the C5 bench config.
mkdir -p gpurun_out
set -o pipefail
{
  echo "== validated suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
  echo "== tests that have only run on tests/simt_emu so far"; timeout 900 python -m pytest tests/test_gpu_zy_late.py tests/test_gpu_zz_options.py -m gpu -q 2>&1 | tail -15
  echo "== ResidLinear on the tcgen05 route (SVAE_RESID_TC=1), A/B on a resid network"
  SVAE_RESID_TC=1 timeout 600 python -m pytest tests/test_gpu_zz_options.py -m gpu -q -k "resid or opt_all" 2>&1 | tail -4
  echo "== CTF fast kernel parity"; SVAE_CTF_FAST=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ctf" 2>&1 | tail -5
  echo "== C5 bench A (default CTF kernel) / B (SVAE_CTF_FAST=1)"
  for v in 0 1; do
    SVAE_CTF_FAST=$v timeout 300 python bench.py --config c5 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ctf_fast=$v', round(d['ms_per_step'],3), 'ms/step', round(d['value']), d['unit'])"
  done
  echo "== C2: dual-stream chunk schedule (SVAE_DUAL_STREAM=1) against the default, eager and graphed"
  for v in 0 1; do for g in "" "--no-graph"; do
    SVAE_DUAL_STREAM=$v timeout 300 python bench.py --config c2 --steps 100 --warmup 5 --no-cpu-baseline $g 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dual_stream=$v $g', round(d['ms_per_step'],3), 'ms/step')"
  done; done
  echo "== C2: images per decoder pass (L2 residency: one (rows x 512) bf16 matrix is 0.8 MB per image; 126 MB L2)"
  # 24 / 48 / 96 images = 18816 / 37632 / 75264 rows = 148 / 294 / 588 pair tiles = 2 / ~4 / ~8 full waves of 74 CTA pairs
  for ch in 0 512 256 96 48 24; do
    timeout 300 python bench.py --config c2 --chunk $ch --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk=$ch', round(d['ms_per_step'],3), 'ms/step', round(d['value']), d['unit'])"
  done
  echo "== C2: small chunks AND the dual-stream schedule"
  for ch in 96 48 24; do
    SVAE_DUAL_STREAM=1 timeout 300 python bench.py --config c2 --chunk $ch --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dual_stream chunk=$ch', round(d['ms_per_step'],3), 'ms/step', round(d['value']), d['unit'])"
  done
} | tee gpurun_out/round2_first.log
