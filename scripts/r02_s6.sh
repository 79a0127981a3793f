#!/bin/bash
# round 2, GPU session 6: cross-attention tcgen05 kernel, packed-fp32x2 softmax, tanh-form GEGLU
mkdir -p gpurun_out
rc=0
run() { name=$1; shift; timeout -k 5 "$TO" "$@" > gpurun_out/s6_$name.log 2>&1; r=$?; echo "== $name rc=$r"; tail -n ${TAILN:-4} gpurun_out/s6_$name.log; return $r; }
PT="python -m pytest -q -m gpu --timeout 240 --timeout-method=thread"
TO=300 run k_attn $PT tests/test_gpu_kernels.py -k "attention or geglu" || rc=1
timeout 200 python scripts/time_attention.py > gpurun_out/s6_attention_times.txt 2>&1; cat gpurun_out/s6_attention_times.txt
RDEIC_B200_LIB=$PWD/rdeic_b200/_build/librdeic_nopoly.so timeout 200 python scripts/time_attention.py 2>&1 | head -8
timeout 300 python scripts/ab_gemm.py 2>&1 | grep -i geglu
RDEIC_B200_LIB=$PWD/rdeic_b200/_build/librdeic_nopoly.so timeout 300 python scripts/ab_gemm.py 2>&1 | grep -i geglu
TO=900 run engine $PT -s tests/test_gpu_engine.py -k "seeds or golden or baseline_latent" || rc=1
grep -E "rel-L2|PSNR" gpurun_out/s6_engine.log | head -30
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err || rc=1
python -c "import json;d=json.load(open('gpurun_out/s6_bench.json'));print('c2:',d['value'],d['unet_step_ms'],d['vae_decode_ms'],d['e2e']['value'],d['roofline']['frac'])"
exit $rc
