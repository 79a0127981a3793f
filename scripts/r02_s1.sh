#!/bin/bash
# round 2, GPU session 1: new conv modes + fp32 entropy nets + engine parity, then bench and the tile sweep
mkdir -p gpurun_out
rc=0
run() { name=$1; shift; timeout -k 5 "$TO" "$@" > gpurun_out/s1_$name.log 2>&1; r=$?; echo "== $name rc=$r"; tail -n ${TAILN:-6} gpurun_out/s1_$name.log; return $r; }
PT="python -m pytest -q -m gpu --timeout 240 --timeout-method=thread"
TO=200 run k_stride2 $PT tests/test_gpu_kernels.py -k "stride2" || { export RDEIC_S2_IM2COL=1; echo "!! stride2 failed -> im2col fallback for the rest"; rc=1; }
TO=200 run k_up2 $PT tests/test_gpu_kernels.py -k "folded" || rc=1
TO=200 run k_inj $PT tests/test_gpu_kernels.py -k "injection" || rc=1
TO=600 run k_rest $PT tests/test_gpu_kernels.py -k "not stride2 and not folded and not injection" || rc=1
TO=600 TAILN=25 run compression $PT tests/test_gpu_compression.py tests/test_gpu_entropy_frontend.py || rc=1
TO=900 TAILN=25 run engine $PT tests/test_gpu_engine.py || rc=1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err || rc=1
cat gpurun_out/s1_bench.json
RDEIC_NO_FUSED_INJECT=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s1_bench_noinj.json 2> gpurun_out/s1_bench_noinj.err
cat gpurun_out/s1_bench_noinj.json
timeout 600 python scripts/tile_sweep.py unet 8 > gpurun_out/s1_tile_sweep_unet.txt 2>&1
head -30 gpurun_out/s1_tile_sweep_unet.txt
timeout 300 python bench.py --compressor 1 > gpurun_out/s1_compressor_b1.json 2> gpurun_out/s1_compressor.err; timeout 300 python bench.py --compressor 1 --compressor-precision bf16 >> gpurun_out/s1_compressor_b1.json 2>> gpurun_out/s1_compressor.err; timeout 300 python bench.py --compressor 8 >> gpurun_out/s1_compressor_b1.json 2>> gpurun_out/s1_compressor.err
cat gpurun_out/s1_compressor_b1.json
exit $rc
