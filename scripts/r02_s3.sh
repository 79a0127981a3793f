#!/bin/bash
# round 2, GPU session 3: hi/lo inputs + fp32 conv1->GN hand-off (parity margin), A/B vs the round-1 library, ncu of two kernels
mkdir -p gpurun_out
rc=0
run() { name=$1; shift; timeout -k 5 "$TO" "$@" > gpurun_out/s3_$name.log 2>&1; r=$?; echo "== $name rc=$r"; tail -n ${TAILN:-4} gpurun_out/s3_$name.log; return $r; }
PT="python -m pytest -q -m gpu --timeout 240 --timeout-method=thread"
TO=300 run k_new $PT tests/test_gpu_kernels.py -k "split_hilo or blend_tiles or attention" || rc=1
TO=900 run seeds $PT -s tests/test_gpu_engine.py -k "seeds or golden or baseline_latent" || rc=1
grep -E "rel-L2|PSNR" gpurun_out/s3_seeds.log
RDEIC_INPUT_HILO=0 TO=900 run seeds_nohilo $PT -s tests/test_gpu_engine.py -k "seeds or full_unet_step_vs"
grep -E "rel-L2|PSNR" gpurun_out/s3_seeds_nohilo.log
timeout 300 python tests/tools/diag_unet_error.py 3 64 > gpurun_out/s3_diag_seed3.txt 2>&1; cat gpurun_out/s3_diag_seed3.txt
RDEIC_B200_LIB=$PWD/rdeic_b200/_build/librdeic_r1.so timeout 300 python scripts/ab_gemm.py > gpurun_out/s3_ab_gemm_r1.txt 2>&1
timeout 300 python scripts/ab_gemm.py > gpurun_out/s3_ab_gemm_new.txt 2>&1
paste -d'|' gpurun_out/s3_ab_gemm_r1.txt gpurun_out/s3_ab_gemm_new.txt | sed -E 's/\|[^ ]+ +/ | /' | cut -c1-160
timeout 200 python scripts/time_attention.py > gpurun_out/s3_attention_times.txt 2>&1; head -3 gpurun_out/s3_attention_times.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err || rc=1
cut -c1-330 gpurun_out/s3_bench.json; python -c "import json;d=json.load(open('gpurun_out/s3_bench.json'));print(d['unet_step_ms'],d['vae_decode_ms'],d['e2e']['value'],d['roofline']['frac'])"
RDEIC_RES_F32=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s3_bench_nores32.json 2> gpurun_out/s3_bench_nores32.err
python -c "import json;d=json.load(open('gpurun_out/s3_bench_nores32.json'));print('RES_F32=0:',d['value'],d['unet_step_ms'],d['vae_decode_ms'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_tc2|conv_gemm" --launch-skip 6 --launch-count 3 -o gpurun_out/s3_two_kernels python scripts/ncu_two_kernels.py > gpurun_out/s3_ncu.log 2>&1 || echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
exit $rc
