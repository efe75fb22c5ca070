#!/bin/bash
for r in 64 112 128 196 256; do
  SVAE_L0_ROWS=$r python bench.py --config c2 --steps 200 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('L0_ROWS=$r', round(d['ms_per_step'],4), 'ms/step; decoder fwd', round(d['decoder_forward']['ms_per_call'],4))"
done
