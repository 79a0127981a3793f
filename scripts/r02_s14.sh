#!/bin/bash
# round 2, GPU session 14: conv_gemm with two MMA-issuing warps (N = 160 tiles, two M tiles per item)
mkdir -p gpurun_out
rc=0
run() { name=$1; shift; timeout -k 5 "$TO" "$@" > gpurun_out/s14_$name.log 2>&1; r=$?; echo "== $name rc=$r"; tail -n ${TAILN:-4} gpurun_out/s14_$name.log; return $r; }
PT="python -m pytest -q -m gpu --timeout 240 --timeout-method=thread"
TO=600 run kernels $PT tests/test_gpu_kernels.py -k "conv or linear or split_k or attention_wide" || rc=1
for v in 0 1; do
  export RDEIC_DUAL160=$v
  echo "== RDEIC_DUAL160=$v"
  timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
  timeout 300 python scripts/gemm_shapes.py unet 8 > gpurun_out/s14_shapes_dual$v.txt 2>&1
  head -12 gpurun_out/s14_shapes_dual$v.txt
done
unset RDEIC_DUAL160
TO=900 run engine $PT -s tests/test_gpu_engine.py -k "golden or baseline_latent or seeds" || rc=1
grep -E "rel-L2|PSNR" gpurun_out/s14_engine.log | head -30
exit $rc
