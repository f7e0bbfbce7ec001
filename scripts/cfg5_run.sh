#!/bin/bash
# BASELINE config 5 on 8 GPUs: N = 1e7 structures x M = 5000 observables (400 GB fp64), N-sharded, 50 GB per GPU
set -u
OUT=gpurun_out/scale
mkdir -p $OUT
for meth in logw forces; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus 8 --observables 5000 --structures 1250000 --method $meth --steps 10 --warmup 3 --no-cpu-baseline \
    > $OUT/cfg5_${meth}_n8.json 2> $OUT/cfg5_${meth}_n8.err
  echo "cfg5 $meth rc=$?"; tail -2 $OUT/cfg5_${meth}_n8.err | cut -c1-300; tail -1 $OUT/cfg5_${meth}_n8.json | cut -c1-1500
done
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -2
