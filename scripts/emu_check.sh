#!/bin/bash
# Before spending GPU minutes on a kernel change: run it on the host models (tests/simt_emu).
#   bash scripts/emu_check.sh            # GEMM unit tests on 148-SM and 4-SM emulated devices + the emulated step suites
#   SVAE_EMU_ASAN=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 bash scripts/emu_check.sh
set -e
cd "$(dirname "$0")/.."
for sms in 148 4; do
  echo "== tcgen05 GEMM unit tests, $sms emulated SMs"
  SVAE_EMU_SMS=$sms SVAE_TEST_BACKEND=emu python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "tc_ and not 78400"
done
echo "== emulated step / fuzz suites, threads scheduled by index and in two pseudo-random orders"
python -m pytest tests/test_emu_step.py tests/test_emu_fuzz.py -q -x -p no:cacheprovider
for seed in 5 9; do
  SVAE_EMU_SHUFFLE=$seed SVAE_EMU_SMS=4 python -m pytest tests/test_emu_step.py tests/test_emu_fuzz.py -q -x -p no:cacheprovider
done
echo "== bench.py main path on the emulation (about 1.5 minutes)"
python tests/emu_bench_smoke.py
