#!/bin/bash
# microbenchmarks (MUFU / conversion / tcgen05 issue rates) + work-item order A/B of the tcgen05 conv kernel
mkdir -p gpurun_out
timeout 120 scripts/micro/ubench all > gpurun_out/s11_ubench.txt 2>&1
timeout 60 scripts/micro/ubench mixed >> gpurun_out/s11_ubench.txt 2>&1
cat gpurun_out/s11_ubench.txt
for mm in 0 1 2; do
  export RDEIC_M_MAJOR=$mm
  echo "== RDEIC_M_MAJOR=$mm"
  timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
  timeout 300 python scripts/gemm_shapes.py unet 8 > gpurun_out/s11_shapes_mm$mm.txt 2>&1
done 2>&1 | tee gpurun_out/s11_mmajor_ab.txt
