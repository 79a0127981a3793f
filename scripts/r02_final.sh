#!/bin/bash
# Round-2 evidence in one call: GPU tests, smoke(), the bench line (with the CPU baseline), the reference arm, the ncu launch
# list of the SAME bench command restricted to its timed device region, the bench line again with roofline.traffic taken
# from that capture, HBM-kernel rooflines, and ncu --set full of one representative launch per kernel class.
mkdir -p gpurun_out
rc=0
timeout -k 5 1200 python -m pytest tests -q -m gpu -s --timeout 600 --timeout-method=thread > gpurun_out/r02_gpu_tests.log 2>&1 || rc=$?
tail -n 2 gpurun_out/r02_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1 || rc=$?
tail -n 1 gpurun_out/r02_smoke.log
timeout 600 python bench.py > gpurun_out/r02_bench_n1_first.json 2> gpurun_out/r02_bench_n1.err || rc=$?
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err || rc=$?
cut -c1-400 gpurun_out/r02_bench_reference_arm.json
timeout 300 python scripts/bench_hbm_kernels.py > gpurun_out/r02_hbm_kernels.json 2> gpurun_out/r02_hbm.err || echo "hbm rc=$?"
timeout 300 python scripts/prof_step.py 8 > gpurun_out/r02_step_timeline.txt 2>&1 || echo "prof rc=$?"
# the report itself (> 64 MiB with source pages) stays on the box: gpurun copies back at most 64 MiB, and a directory over the limit is dropped whole
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o /tmp/r02_kernels_full python scripts/profile_final.py > gpurun_out/r02_ncu_full.log 2>&1 || echo "ncu full rc=$?"
python scripts/ncu_summary.py /tmp/r02_kernels_full.ncu-rep > gpurun_out/r02_kernels_ncu_full_summary.txt 2>/dev/null
ncu -i /tmp/r02_kernels_full.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02_kernels_ncu_full_raw.csv.gz
timeout ${NCU_LIMIT:-1000} ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --cache-control none --profile-from-start off --csv --log-file gpurun_out/r02_bench_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-range > gpurun_out/r02_bench_ncu.log 2>&1 || echo "ncu list rc=$?"
python scripts/kernel_summary.py gpurun_out/r02_bench_launches.csv gpurun_out/r02_bench_kernel_summary.json \
  "ncu --metrics gpu__time_duration.sum,dram__bytes_{read,write}.sum --clock-control none --cache-control none --profile-from-start off python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-range (the timed device region of one 512^2 batch-8 5-step decode, CUDA-graph kernel nodes)"
python scripts/summarize_launches.py gpurun_out/r02_bench_launches.csv --grid > gpurun_out/r02_bench_launches_by_grid.txt 2>/dev/null || true
gzip -f gpurun_out/r02_bench_launches.csv
timeout 600 python bench.py --traffic-from gpurun_out/r02_bench_kernel_summary.json > gpurun_out/r02_bench_n1.json 2>> gpurun_out/r02_bench_n1.err || rc=$?
cat gpurun_out/r02_bench_n1.json
exit $rc
