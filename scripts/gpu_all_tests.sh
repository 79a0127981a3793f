#!/bin/bash
mkdir -p gpurun_out
rc=0
timeout -k 5 1200 python -m pytest tests -q -m gpu -s --timeout 600 --timeout-method=thread > gpurun_out/gpu_tests.log 2>&1 || rc=$?
echo "== gpu tests rc=$rc"; grep -E "PSNR|rel-L2|passed|failed|Error" gpurun_out/gpu_tests.log | tail -n 30
exit $rc
