#!/bin/bash
# vm kernels, second generation: parity tests, then the -vm bench line per variant
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_field.py -m gpu -x -q -k "vm or preset" > gpurun_out/r2v_tests.log 2>&1; tail -3 gpurun_out/r2v_tests.log
for v in "field_planes_v2=0" "field_planes_v2=1" "field_planes_v2=1,field_planes_unroll=1" "field_planes_v2=1,field_planes_unroll=2" "field_planes_v2=1,field_planes_unroll=3"; do
  FFB_TUNING=$v timeout 300 python bench.py --workload nerf_vm --no-cpu-baseline --no-cuda-eager-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
  python - "$v" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['ms_per_step'],4), {k:v['ms_per_step'] for k,v in d['kernels'].items()})
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open('gpurun_out/r2v_bench.err').read()[-1500:])
PY
done
