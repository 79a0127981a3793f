#!/bin/bash
# round 2, GPU session 17: N = 160 tile pairs sharing one activation stage (kNT = 2); GroupNorm apply as one resident wave
mkdir -p gpurun_out
rc=0
run() { name=$1; shift; timeout -k 5 "$TO" "$@" > gpurun_out/s17_$name.log 2>&1; r=$?; echo "== $name rc=$r"; tail -n ${TAILN:-4} gpurun_out/s17_$name.log; return $r; }
PT="python -m pytest -q -m gpu --timeout 240 --timeout-method=thread"
TO=900 run kernels $PT tests/test_gpu_kernels.py || rc=1
for v in 0 1; do
  export RDEIC_PAIR160=$v
  echo "== RDEIC_PAIR160=$v"
  timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
  timeout 300 python scripts/gemm_shapes.py unet 8 > gpurun_out/s17_shapes_pair$v.txt 2>&1
  head -12 gpurun_out/s17_shapes_pair$v.txt
done
unset RDEIC_PAIR160
timeout 300 python scripts/prof_step.py 8 > gpurun_out/s17_step_timeline.txt 2>&1; grep -E "gn_|span" gpurun_out/s17_step_timeline.txt
TO=900 run engine $PT -s tests/test_gpu_engine.py || rc=1
grep -E "rel-L2|PSNR" gpurun_out/s17_engine.log | head -40
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s17_bench.json 2> gpurun_out/s17_bench.err || rc=1
python -c "import json;d=json.load(open('gpurun_out/s17_bench.json'));print('c2:',d['value'],d['unet_step_ms'],d['vae_decode_ms'],d['e2e']['value'],d['roofline']['frac'])"
exit $rc
