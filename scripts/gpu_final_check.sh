#!/bin/bash
# Round-end evidence in one call: GPU tests, smoke(), the bench line (about 3 GPU-minutes), and -- only with
# the argument `ncu` -- the ncu launch list of the SAME bench command restricted to its timed device region
# (--profile-range + --profile-from-start off).  ncu profiles CUDA-graph kernel nodes at ~0.18 s each: the
# ~3 160 launches of one timed decode need about 10 GPU-minutes, so the capture is bounded by `timeout`.
mkdir -p gpurun_out
rc=0
timeout -k 5 900 python -m pytest tests -q -m gpu --timeout 600 --timeout-method=thread > gpurun_out/gpu_tests.log 2>&1 || rc=$?
tail -n 2 gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || rc=$?
tail -n 1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/r01_bench_n1.json 2> gpurun_out/r01_bench_n1.err || rc=$?
cat gpurun_out/r01_bench_n1.json
if [ $rc -eq 0 ] && [ "$1" = "ncu" ]; then
  timeout ${NCU_LIMIT:-600} ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --cache-control none --profile-from-start off --csv --log-file gpurun_out/r01_bench_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-range > gpurun_out/bench_ncu.log 2>&1 || echo "ncu rc=$?"
  python scripts/kernel_summary.py gpurun_out/r01_bench_launches.csv gpurun_out/r01_bench_kernel_summary.json \
    "ncu --metrics gpu__time_duration.sum,dram__bytes_{read,write}.sum --clock-control none --cache-control none --profile-from-start off python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-range (the timed device region of one 512^2 batch-8 5-step decode, CUDA-graph kernel nodes)"
  python scripts/summarize_launches.py gpurun_out/r01_bench_launches.csv --grid > gpurun_out/r01_bench_launches_by_grid.txt 2>/dev/null || true
  rm -f gpurun_out/r01_bench_launches.csv.tmp
fi
exit $rc
