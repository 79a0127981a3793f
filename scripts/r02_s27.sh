#!/bin/bash
# round 2, GPU session 27: one-kernel GroupNorm with the group's elements held in registers (24 / 40 pairs per thread) -- tests, then A/B on one box
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu --timeout 240 --timeout-method=thread"
timeout 600 $PT tests/test_gpu_kernels.py -k "groupnorm" 2>&1 | tail -2
timeout 900 $PT -s tests/test_gpu_engine.py -k "golden or seeds" 2>&1 | grep -E "passed|failed|rel-L2" | head -12
for rep in 1 2 3; do
  for v in prev cur; do
    if [ $v = cur ]; then unset RDEIC_B200_LIB; else export RDEIC_B200_LIB=$PWD/rdeic_b200/_build/librdeic_$v.so; fi
    echo "== lib=$v (rep $rep)"
    timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
  done
done 2>&1 | tee gpurun_out/s27_gn_small_ab.txt
unset RDEIC_B200_LIB
timeout 300 python scripts/prof_step.py 8 2>&1 | grep -E "gn_|span"
