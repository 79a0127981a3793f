#!/bin/bash
# round 2, GPU session 26: PDL with the implicit trigger everywhere vs an entry trigger in the short kernels only
mkdir -p gpurun_out
for rep in 1 2 3; do
  for v in cur shortearly; do
    if [ $v = cur ]; then unset RDEIC_B200_LIB; else export RDEIC_B200_LIB=$PWD/rdeic_b200/_build/librdeic_$v.so; fi
    echo "== lib=$v (rep $rep)"
    timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
  done
done 2>&1 | tee gpurun_out/s26_pdl_short_ab.txt
