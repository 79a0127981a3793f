#!/bin/bash
# round 2, GPU session 22: A/B of the conv kernel's warp layout (320 threads vs control warp group + setmaxnreg) on one box
mkdir -p gpurun_out
for rep in 1 2; do
for v in prev cur; do
  if [ $v = cur ]; then unset RDEIC_B200_LIB; else export RDEIC_B200_LIB=$PWD/rdeic_b200/_build/librdeic_prev.so; fi
  echo "== $v (rep $rep)"
  timeout 300 python scripts/gemm_shapes.py vae 8 2>&1 | head -8
  timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
done
done 2>&1 | tee gpurun_out/s22_layout_ab.txt
