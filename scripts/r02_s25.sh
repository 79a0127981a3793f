#!/bin/bash
# round 2, GPU session 25: programmatic dependent launch on by default (implicit trigger): the whole GPU suite, smoke, the bench, config 4, the compressor
mkdir -p gpurun_out
rc=0
PT="python -m pytest -q -m gpu --timeout 600 --timeout-method=thread"
timeout 1200 $PT tests 2>&1 | tail -3 || rc=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 || rc=1
for pdl in 0 1; do
  export RDEIC_PDL=$pdl
  echo "== RDEIC_PDL=$pdl"
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s25_bench_pdl$pdl.json 2> gpurun_out/s25_bench_pdl$pdl.err || rc=1
  python -c "import json;d=json.load(open('gpurun_out/s25_bench_pdl$pdl.json'));print('c2:',d['value'],d['unet_step_ms'],d['vae_decode_ms'],d['e2e']['value'],d['roofline']['frac'])"
  timeout 600 python bench.py --config c4 --steps 3 --warmup 3 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('c4:',d['ms_per_step'],d['tiles'])"
  timeout 600 python bench.py --compressor 1 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('compressor:',d['compress_ms'],d['decompress_ms'])"
done
exit $rc
