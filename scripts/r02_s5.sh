#!/bin/bash
# round 2, GPU session 5: attention exp2 variants A/B on one box; compute-sanitizer over the kernel tests
mkdir -p gpurun_out
B=$PWD/rdeic_b200/_build
for rep in 1 2; do
  for v in nopoly default poly2; do
    if [ $v = default ]; then unset RDEIC_B200_LIB; else export RDEIC_B200_LIB=$B/librdeic_$v.so; fi
    echo "== $v (rep $rep)"; timeout 200 python scripts/time_attention.py 2>&1 | head -2
  done
done > gpurun_out/s5_attention_ab.txt 2>&1
unset RDEIC_B200_LIB
cat gpurun_out/s5_attention_ab.txt
PT="python -m pytest -q -m gpu -x --timeout 1200 --timeout-method=thread -p no:cacheprovider"
timeout -k 5 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 $PT tests/test_gpu_kernels.py > gpurun_out/s5_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 6 gpurun_out/s5_sanitizer_memcheck.log
timeout -k 5 600 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 $PT tests/test_gpu_kernels.py -k "conv3x3_tc or attention or groupnorm or linear_tc or layernorm or fused_phase or vq_ or tail or split_k" > gpurun_out/s5_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 6 gpurun_out/s5_sanitizer_racecheck.log
