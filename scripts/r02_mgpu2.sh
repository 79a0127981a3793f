#!/bin/bash
# 2-GPU bring-up of the strong-scaling configs (cheap) before the 8-GPU run
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511 --nproc-per-node 2"
for c in c3 c4 c5; do
  timeout 600 $T bench.py --config $c --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_${c}_n2.json 2> gpurun_out/r02_${c}_n2.err || tail -8 gpurun_out/r02_${c}_n2.err
  cut -c1-900 gpurun_out/r02_${c}_n2.json
done
