#!/bin/bash
# N-GPU bench under torchrun with a watchdog: N=${N:-2}; extra flags in $FLAGS
mkdir -p gpurun_out
N=${N:-2}
export NCCL_DEBUG=${NCCL_DEBUG:-WARN}
timeout ${WD:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps ${STEPS:-100} --warmup 5 --no-cpu-baseline $FLAGS > gpurun_out/multi_$N.json 2> gpurun_out/multi_$N.err
echo "rc=$?"
tail -c 1500 gpurun_out/multi_$N.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/multi_$N.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","host_enqueue_ms_per_step")}, d["config"]["launch"])
    for k,v in d.get("extra_configs",{}).items(): print(k, v)
except Exception as e: print("no json", e)
PY
