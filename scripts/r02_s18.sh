#!/bin/bash
# round 2, GPU session 18: GroupNorm apply (raw prefetch, one resident wave) timings; config 3 on one GPU with the flash VAE attention
mkdir -p gpurun_out
rc=0
PT="python -m pytest -q -m gpu --timeout 240 --timeout-method=thread"
timeout 300 $PT tests/test_gpu_kernels.py -k "groupnorm or norm" 2>&1 | tail -2
timeout 200 python scripts/time_gn_fused.py 2>&1 | tee gpurun_out/s18_gn_fused.txt
timeout 200 python scripts/time_gn.py 2>&1 | tee gpurun_out/s18_gn.txt
timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
timeout 900 python bench.py --config c3 --steps 3 --warmup 3 > gpurun_out/s18_bench_c3.json 2> gpurun_out/s18_bench_c3.err || rc=1
python -c "import json;d=json.load(open('gpurun_out/s18_bench_c3.json'));print('c3:',d['value'],d['ms_per_step'],d.get('peak_mem_gib'),d.get('e2e'))"
exit $rc
