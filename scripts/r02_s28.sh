#!/bin/bash
# round 2, GPU session 28: do the tile-picker / epilogue-width constants still sit at their optimum after the issue-rate fix?
mkdir -p gpurun_out
{
echo "== defaults"; timeout 200 python scripts/ab_unet.py 8 2>&1 | tail -1
for v in 20 45; do echo "== RDEIC_EPI12_KB=$v"; RDEIC_EPI12_KB=$v timeout 200 python scripts/ab_unet.py 8 2>&1 | tail -1; done
for v in 48 144; do echo "== RDEIC_TILE_OVERHEAD=$v"; RDEIC_TILE_OVERHEAD=$v timeout 200 python scripts/ab_unet.py 8 2>&1 | tail -1; done
echo "== defaults"; timeout 200 python scripts/ab_unet.py 8 2>&1 | tail -1
} 2>&1 | tee gpurun_out/s28_constants.txt
