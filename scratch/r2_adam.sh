#!/bin/bash
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT"
for w in nerf sdf image; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --no-cuda-eager-baseline > gpurun_out/r2a_bench_$w.json 2> gpurun_out/r2a_bench_$w.err
  python - $w <<'PY'
import json,sys
w=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r2a_bench_{w}.json').read().strip().splitlines()[-1])
    print(w, round(d['ms_per_step'],4), {k:v['ms_per_step'] for k,v in d['kernels'].items()})
except Exception as e:
    print(w, 'FAILED', e); print(open(f'gpurun_out/r2a_bench_{w}.err').read()[-1500:])
PY
done
