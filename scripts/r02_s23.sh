#!/bin/bash
# round 2, GPU session 23: residual loads hoisted ahead of the accumulator read in the conv epilogue -- A/B against the previous build on one box
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu --timeout 240 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_kernels.py -k "conv or linear or split_k or groupnorm_from" 2>&1 | tail -2
for rep in 1 2; do
for v in prev cur; do
  if [ $v = cur ]; then unset RDEIC_B200_LIB; else export RDEIC_B200_LIB=$PWD/rdeic_b200/_build/librdeic_prev.so; fi
  echo "== $v (rep $rep)"
  timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
  if [ $rep = 1 ]; then
    timeout 300 python scripts/gemm_shapes.py unet 8 2>&1 | head -22
    timeout 300 python scripts/gemm_shapes.py vae 8 2>&1 | head -8
  fi
done
done 2>&1 | tee gpurun_out/s23_resid_ab.txt
