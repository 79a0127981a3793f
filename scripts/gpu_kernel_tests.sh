#!/bin/bash
# Run the GPU kernel parity tests group by group, each under its own timeout so a hung
# kernel cannot take the whole call (or the box) with it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for k in "ckbd or quantize or build_indexes or fused_phase or vq" "relay or q_sample or layout or timestep or geglu or image" "groupnorm or layernorm" "linear_tc or conv" "attention"; do
  name=$(echo "$k" | cut -d' ' -f1)
  timeout -k 5 420 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$k" -x --timeout 180 --timeout-method=thread > gpurun_out/k_$name.log 2>&1
  r=$?
  echo "== group [$k] rc=$r"; tail -n 25 gpurun_out/k_$name.log
  [ $r -ne 0 ] && rc=$r
done
exit $rc
