#!/bin/bash
# round 2, GPU session 8: single-kernel GroupNorm for small tensors, post-upsample injections folded into the up2 conv
mkdir -p gpurun_out
rc=0
run() { name=$1; shift; timeout -k 5 "$TO" "$@" > gpurun_out/s8_$name.log 2>&1; r=$?; echo "== $name rc=$r"; tail -n ${TAILN:-4} gpurun_out/s8_$name.log; return $r; }
PT="python -m pytest -q -m gpu --timeout 240 --timeout-method=thread"
TO=600 run kernels $PT tests/test_gpu_kernels.py || rc=1
TO=900 run engine $PT -s tests/test_gpu_engine.py || rc=1
grep -E "rel-L2|PSNR" gpurun_out/s8_engine.log | head -44
TO=600 run compression $PT tests/test_gpu_compression.py tests/test_gpu_entropy_frontend.py || rc=1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err || rc=1
python -c "import json;d=json.load(open('gpurun_out/s8_bench.json'));print('c2:',d['value'],d['unet_step_ms'],d['vae_decode_ms'],d['e2e']['value'],d['roofline']['frac'],d['gpu_launches'])"
exit $rc
