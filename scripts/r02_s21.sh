#!/bin/bash
# round 2, GPU session 21: conv_gemm with a control warp group + setmaxnreg (epilogue warps 224 / 152 registers); fp32 residual prefetch A/B
mkdir -p gpurun_out
rc=0
PT="python -m pytest -q -m gpu --timeout 240 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_kernels.py -k "conv or linear or split_k or groupnorm or tail" 2>&1 | tail -3 || rc=1
for v in 0 1; do
  if [ $v = 1 ]; then export RDEIC_RESID_AHEAD=1; else unset RDEIC_RESID_AHEAD; fi
  echo "== RDEIC_RESID_AHEAD=$v"
  timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
  timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
  timeout 300 python scripts/gemm_shapes.py unet 8 > gpurun_out/s21_shapes_ahead$v.txt 2>&1
  head -16 gpurun_out/s21_shapes_ahead$v.txt
done
unset RDEIC_RESID_AHEAD
timeout 300 python scripts/gemm_shapes.py vae 8 > gpurun_out/s21_shapes_vae.txt 2>&1; head -8 gpurun_out/s21_shapes_vae.txt
exit $rc
