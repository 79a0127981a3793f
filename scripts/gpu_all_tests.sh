#!/bin/bash
mkdir -p gpurun_out
rc=0
timeout -k 5 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 --timeout-method=thread > gpurun_out/kernels.log 2>&1 || rc=$?
echo "== kernels rc=$rc"; tail -n 12 gpurun_out/kernels.log
timeout -k 5 900 python -m pytest tests/test_gpu_engine.py -q -m gpu -s --timeout 600 --timeout-method=thread > gpurun_out/engine.log 2>&1 || rc=$?
echo "== engine rc=$rc"; grep -E "^\.*F*\[|PSNR|rel-L2|passed|failed" gpurun_out/engine.log | tail -n 40
exit $rc
