#!/bin/bash
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT"
for v in "field_level_parallel=1" "field_level_parallel=0" "field_level_parallel=0,field_fwd_stage=0"; do
for w in image sdf image_set; do
  FFB_TUNING=$v timeout 300 python bench.py --workload $w --no-cpu-baseline --no-cuda-eager-baseline > gpurun_out/r2s_bench_$w.json 2> gpurun_out/r2s_bench_$w.err
  python - $w $v <<'PY'
import json,sys
w=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r2s_bench_{w}.json').read().strip().splitlines()[-1])
    print(sys.argv[2], w, round(d['ms_per_step'],4), {k:v['ms_per_step'] for k,v in d['kernels'].items()}, d['roofline']['frac'])
except Exception as e:
    print(w, 'FAILED', e); print(open(f'gpurun_out/r2s_bench_{w}.err').read()[-1500:])
PY
done; done
