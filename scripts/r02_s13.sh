#!/bin/bash
# round 2, GPU session 13: two MMA-issuing warps in the attention kernels, flash attention for the VAE mid block (d = 512)
mkdir -p gpurun_out
rc=0
run() { name=$1; shift; timeout -k 5 "$TO" "$@" > gpurun_out/s13_$name.log 2>&1; r=$?; echo "== $name rc=$r"; tail -n ${TAILN:-4} gpurun_out/s13_$name.log; return $r; }
PT="python -m pytest -q -m gpu --timeout 240 --timeout-method=thread"
TO=300 run k_attn $PT tests/test_gpu_kernels.py -k "attention" || rc=1
timeout 200 python scripts/time_attention.py > gpurun_out/s13_attention_times.txt 2>&1; cat gpurun_out/s13_attention_times.txt
TO=900 run engine $PT -s tests/test_gpu_engine.py -k "vae or golden or baseline_latent or full_size" || rc=1
grep -E "rel-L2|PSNR" gpurun_out/s13_engine.log | head -30
timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
exit $rc
