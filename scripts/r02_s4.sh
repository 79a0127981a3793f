#!/bin/bash
# round 2, GPU session 4: poly-exp attention, kUp conv instantiations, the other BASELINE configs at N=1 (bring-up)
mkdir -p gpurun_out
rc=0
run() { name=$1; shift; timeout -k 5 "$TO" "$@" > gpurun_out/s4_$name.log 2>&1; r=$?; echo "== $name rc=$r"; tail -n ${TAILN:-4} gpurun_out/s4_$name.log; return $r; }
PT="python -m pytest -q -m gpu --timeout 240 --timeout-method=thread"
TO=600 run kernels $PT tests/test_gpu_kernels.py || rc=1
timeout 200 python scripts/time_attention.py > gpurun_out/s4_attention_times.txt 2>&1; head -4 gpurun_out/s4_attention_times.txt
RDEIC_B200_LIB=$PWD/rdeic_b200/_build/librdeic_r1.so timeout 300 python scripts/ab_gemm.py > gpurun_out/s4_ab_gemm_r1.txt 2>&1
timeout 300 python scripts/ab_gemm.py > gpurun_out/s4_ab_gemm_new.txt 2>&1
paste -d'|' gpurun_out/s4_ab_gemm_r1.txt gpurun_out/s4_ab_gemm_new.txt | sed -E 's/\|[^ ]+ +/ | /' | cut -c33-160
TO=900 run engine $PT -s tests/test_gpu_engine.py || rc=1
grep -E "rel-L2|PSNR" gpurun_out/s4_engine.log | head -40
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err || rc=1
python -c "import json;d=json.load(open('gpurun_out/s4_bench.json'));print('c2:',d['value'],d['unet_step_ms'],d['vae_decode_ms'],d['e2e']['value'],d['roofline']['frac'])"
for c in c3 c5 c4; do
  timeout 900 python bench.py --config $c --steps 2 --warmup 3 > gpurun_out/s4_bench_$c.json 2> gpurun_out/s4_bench_$c.err || { rc=1; tail -5 gpurun_out/s4_bench_$c.err; }
  cut -c1-1200 gpurun_out/s4_bench_$c.json
done
exit $rc
