#!/bin/bash
# round 2, GPU session 24: programmatic dependent launch once more, with today's kernels: off / on with the trigger at kernel start / on with the implicit trigger at CTA exit
mkdir -p gpurun_out
for rep in 1 2; do
  for v in "0 cur" "1 cur" "1 notrig"; do
    set -- $v
    export RDEIC_PDL=$1
    if [ $2 = cur ]; then unset RDEIC_B200_LIB; else export RDEIC_B200_LIB=$PWD/rdeic_b200/_build/librdeic_$2.so; fi
    echo "== RDEIC_PDL=$1 lib=$2 (rep $rep)"
    timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
  done
done 2>&1 | tee gpurun_out/s24_pdl_ab.txt
