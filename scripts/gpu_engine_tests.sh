#!/bin/bash
mkdir -p gpurun_out
rc=0
timeout -k 5 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "ckbd or quantize or build_indexes or fused_phase or vq" --timeout 120 --timeout-method=thread > gpurun_out/k_entropy.log 2>&1 || rc=$?
echo "== entropy rc=$rc"; tail -n 15 gpurun_out/k_entropy.log
timeout -k 5 900 python -m pytest tests/test_gpu_engine.py -q -m gpu -s --timeout 600 --timeout-method=thread > gpurun_out/engine.log 2>&1 || rc=$?
echo "== engine rc=$rc"; grep -E "^\[|passed|failed|Error|error" gpurun_out/engine.log | tail -n 40; tail -n 30 gpurun_out/engine.log
exit $rc
