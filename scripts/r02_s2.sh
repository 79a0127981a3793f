#!/bin/bash
# round 2, GPU session 2: producer-loop fix + TMEM-resident P attention, parity diagnostics
mkdir -p gpurun_out
rc=0
run() { name=$1; shift; timeout -k 5 "$TO" "$@" > gpurun_out/s2_$name.log 2>&1; r=$?; echo "== $name rc=$r"; tail -n ${TAILN:-6} gpurun_out/s2_$name.log; return $r; }
PT="python -m pytest -q -m gpu --timeout 240 --timeout-method=thread"
TO=300 run k_attn $PT tests/test_gpu_kernels.py -k "attention" || rc=1
TO=300 run k_conv $PT tests/test_gpu_kernels.py -k "conv or linear or split_k or stats" || rc=1
TO=200 python scripts/time_attention.py > gpurun_out/s2_attention_times.txt 2>&1; cat gpurun_out/s2_attention_times.txt
RDEIC_ATTN_LAZY=0 TO=200 python scripts/time_attention.py > gpurun_out/s2_attention_times_nolazy.txt 2>&1; head -3 gpurun_out/s2_attention_times_nolazy.txt
timeout 600 python scripts/tile_sweep.py unet 8 > gpurun_out/s2_tile_sweep_unet.txt 2>&1
head -14 gpurun_out/s2_tile_sweep_unet.txt
TO=900 TAILN=40 run seeds $PT -s tests/test_gpu_engine.py -k "seeds or golden or baseline_latent"
grep -E "rel-L2|PSNR" gpurun_out/s2_seeds.log
RDEIC_RES_F32=1 TO=900 run seeds_resf32 $PT -s tests/test_gpu_engine.py -k "seeds or golden or baseline_latent"
grep -E "rel-L2|PSNR" gpurun_out/s2_seeds_resf32.log
timeout 300 python tests/tools/diag_unet_error.py 3 64 > gpurun_out/s2_diag_seed3.txt 2>&1; cat gpurun_out/s2_diag_seed3.txt
RDEIC_RES_F32=1 timeout 300 python tests/tools/diag_unet_error.py 3 64 > gpurun_out/s2_diag_seed3_resf32.txt 2>&1; cat gpurun_out/s2_diag_seed3_resf32.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err || rc=1
cat gpurun_out/s2_bench.json
RDEIC_NO_FUSED_INJECT=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s2_bench_noinj.json 2> gpurun_out/s2_bench_noinj.err
cat gpurun_out/s2_bench_noinj.json
exit $rc
