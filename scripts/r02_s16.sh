#!/bin/bash
# round 2, GPU session 16: single-thread roles entered through elect.sync (no ELECT / BRA.U.ANY wrapper per tcgen05.mma)
mkdir -p gpurun_out
rc=0
run() { name=$1; shift; timeout -k 5 "$TO" "$@" > gpurun_out/s16_$name.log 2>&1; r=$?; echo "== $name rc=$r"; tail -n ${TAILN:-4} gpurun_out/s16_$name.log; return $r; }
PT="python -m pytest -q -m gpu --timeout 240 --timeout-method=thread"
TO=900 run kernels $PT tests/test_gpu_kernels.py || rc=1
timeout 200 python scripts/time_attention.py > gpurun_out/s16_attention_times.txt 2>&1; cat gpurun_out/s16_attention_times.txt
timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
timeout 300 python scripts/gemm_shapes.py unet 8 > gpurun_out/s16_shapes_unet.txt 2>&1; head -14 gpurun_out/s16_shapes_unet.txt
timeout 300 python scripts/gemm_shapes.py vae 8 > gpurun_out/s16_shapes_vae.txt 2>&1; head -10 gpurun_out/s16_shapes_vae.txt
TO=900 run engine $PT -s tests/test_gpu_engine.py || rc=1
grep -E "rel-L2|PSNR" gpurun_out/s16_engine.log | head -40
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s16_bench.json 2> gpurun_out/s16_bench.err || rc=1
python -c "import json;d=json.load(open('gpurun_out/s16_bench.json'));print('c2:',d['value'],d['unet_step_ms'],d['vae_decode_ms'],d['e2e']['value'],d['roofline']['frac'])"
exit $rc
