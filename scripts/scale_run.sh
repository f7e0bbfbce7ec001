#!/bin/bash
# Multi-GPU runs on one box (gpurun --gpus 8): sharded parity check, then the bench at 1/2/4/8 GPUs (weak scaling,
# N = 1e6 per GPU), the forces method at 8, and BASELINE config 5 (N = 1e7 x M = 5000 over 8 GPUs = 50 GB per GPU).
set -u
OUT=gpurun_out/scale
mkdir -p $OUT
run() {  # run <n> <tag> <bench args...>
  local n=$1 tag=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 "$@" > $OUT/$tag.json 2> $OUT/$tag.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $((29600 + n)) bench.py --gpus $n "$@" > $OUT/$tag.json 2> $OUT/$tag.err
  fi
  echo "$tag rc=$? $(tail -c 400 $OUT/$tag.json | head -c 10 >/dev/null; python - <<PY
import json
try:
    d=json.loads(open("$OUT/$tag.json").read().strip().splitlines()[-1])
    print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "pass GB/s", round(d["roofline"]["achieved"]), "e2e", round(d["e2e"]["value"],1), "t2o", d.get("time_to_optimum",{}).get("seconds"))
except Exception as e:
    print("no json:", e)
PY
)"
}
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -2
for n in 1 2 4 8; do run $n logw_n$n --steps 50 --warmup 5 --no-cpu-baseline; done
run 8 forces_n8 --method forces --steps 50 --warmup 5 --no-cpu-baseline
run 8 cfg5_logw_n8 --observables 5000 --structures 1250000 --steps 10 --warmup 3 --no-cpu-baseline
run 8 cfg5_forces_n8 --observables 5000 --structures 1250000 --method forces --steps 10 --warmup 3 --no-cpu-baseline
nvidia-smi --query-gpu=index,memory.used --format=csv | head -3
