#!/bin/bash
# 2-GPU check of the final bench (config 2, both arms launched the way the driver launches them) + config 3 with the flash VAE attention
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29512 --nproc-per-node 2"
timeout 600 $T bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_c2_n2.json 2> gpurun_out/r02_c2_n2.err || tail -8 gpurun_out/r02_c2_n2.err
cut -c1-600 gpurun_out/r02_c2_n2.json
timeout 600 $T bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r02_c2_n2_reference.json 2> gpurun_out/r02_c2_n2_reference.err || tail -8 gpurun_out/r02_c2_n2_reference.err
cut -c1-400 gpurun_out/r02_c2_n2_reference.json
timeout 600 $T bench.py --config c3 --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_c3_n2b.json 2> gpurun_out/r02_c3_n2b.err || tail -8 gpurun_out/r02_c3_n2b.err
cut -c1-400 gpurun_out/r02_c3_n2b.json
timeout 300 python -m pytest -q -m gpu tests/test_gpu_engine.py -k "tiled or parallel or gather" 2>&1 | tail -2
