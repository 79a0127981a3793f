#!/bin/bash
# round 2, GPU session 19: GroupNorm apply variants A (round-1 kernel), B (fp32 prefetch + one resident wave), C (raw prefetch, launch bounds)
mkdir -p gpurun_out
for v in A B C; do
  if [ $v = C ]; then unset RDEIC_B200_LIB; else export RDEIC_B200_LIB=$PWD/rdeic_b200/_build/librdeic_gn$v.so; fi
  echo "== variant $v"
  timeout 200 python scripts/time_gn_fused.py 2>&1 | sed 's/conv.*groupnorm/groupnorm/'
  timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
  timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
done 2>&1 | tee gpurun_out/s19_gn_ab.txt
