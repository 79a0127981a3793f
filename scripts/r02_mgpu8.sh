#!/bin/bash
# BASELINE configs 3 / 4 / 5 on the 8 GPUs of one box (strong scaling; uint8 gather inside the timed region)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29512"
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err || tail -8 gpurun_out/$name.err; cut -c1-1000 gpurun_out/$name.json; }
run r02_c3_n8 $T --nproc-per-node 8 bench.py --config c3 --gpus 8 --steps 3 --warmup 3
run r02_c3_n4 $T --nproc-per-node 4 bench.py --config c3 --gpus 4 --steps 3 --warmup 3
run r02_c5_n8 $T --nproc-per-node 8 bench.py --config c5 --gpus 8 --steps 4 --warmup 3
run r02_c4_n8 $T --nproc-per-node 8 bench.py --config c4 --gpus 8 --steps 4 --warmup 3
run r02_c4_n8_16tiles $T --nproc-per-node 8 bench.py --config c4 --gpus 8 --steps 4 --warmup 3 --c4-max-tile-area 6400
run r02_c2_n8 $T --nproc-per-node 8 bench.py --gpus 8 --steps 5 --warmup 3
