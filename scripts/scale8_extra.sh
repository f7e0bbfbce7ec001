#!/bin/bash
# 8-GPU extras: strong scaling of the N = 1e6 problem (BASELINE config 3 read literally) and the sharded theta scan
set -u
OUT=gpurun_out/scale
mkdir -p $OUT
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:3}" > $OUT/$2.json 2> $OUT/$2.err; echo "$2 rc=$?"; tail -1 $OUT/$2.json | cut -c1-2000; }
tr 29701 strong_logw_n8 --strong --steps 100 --warmup 5 --no-cpu-baseline
tr 29702 strong_forces_n8 --strong --method forces --steps 100 --warmup 5 --no-cpu-baseline
tr 29703 scan32_logw_n8 --theta-scan 32 --steps 10 --warmup 3
