#!/bin/bash
# the final kernels on the 8 GPUs of one box: config 2 (weak), config 3 / 5 (strong, gather timed), config 4 (one frame as tiles)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29513"
run() { name=$1; shift; timeout 400 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err || tail -8 gpurun_out/$name.err; cut -c1-330 gpurun_out/$name.json; }
run r02b_c2_n8 $T --nproc-per-node 8 bench.py --gpus 8 --steps 5 --warmup 3
run r02b_c3_n8 $T --nproc-per-node 8 bench.py --config c3 --gpus 8 --steps 3 --warmup 3
run r02b_c5_n8 $T --nproc-per-node 8 bench.py --config c5 --gpus 8 --steps 4 --warmup 3
run r02b_c4_n8 $T --nproc-per-node 8 bench.py --config c4 --gpus 8 --steps 4 --warmup 3 --c4-max-tile-area 6400
