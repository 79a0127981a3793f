#!/bin/bash
mkdir -p gpurun_out
{
echo "== default";            timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
echo "== RDEIC_NO_OVERLAP=1"; RDEIC_NO_OVERLAP=1 timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
echo "== RDEIC_NO_PDL=1";     RDEIC_NO_PDL=1 timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
echo "== RDEIC_NO_PDL=1 RDEIC_NO_OVERLAP=1"; RDEIC_NO_PDL=1 RDEIC_NO_OVERLAP=1 timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
echo "== default again";      timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
} > gpurun_out/s9_ab_step.txt 2>&1
cat gpurun_out/s9_ab_step.txt
timeout 300 python -m pytest -q -m gpu tests/test_gpu_kernels.py -k groupnorm 2>&1 | tail -2
