#!/bin/bash
mkdir -p gpurun_out
for pdl in 0 1; do
  export RDEIC_PDL=$pdl
  echo "== RDEIC_PDL=$pdl"
  timeout 300 python scripts/ab_unet.py 8 2>&1 | tail -1
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s10_bench_pdl$pdl.json 2> gpurun_out/s10_bench_pdl$pdl.err
  python -c "import json;d=json.load(open('gpurun_out/s10_bench_pdl$pdl.json'));print('c2:',d['value'],d['unet_step_ms'],d['vae_decode_ms'],d['e2e']['value'],d['roofline']['frac'])"
  timeout 600 python bench.py --config c4 --steps 3 --warmup 3 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('c4:',d['ms_per_step'],d['tiles'])"
  timeout 600 python bench.py --compressor 1 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('compressor:',d['compress_ms'],d['decompress_ms'])"
done 2>&1 | tee gpurun_out/s10_pdl_ab.txt
