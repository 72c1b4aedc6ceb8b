#!/bin/bash
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT"
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-cuda-eager-baseline > gpurun_out/r2g_bench1.json 2> gpurun_out/r2g_bench1.err
timeout 300 python bench.py --workload nerf_eval > gpurun_out/r2f_bench_nerf_eval.json 2> gpurun_out/r2f_bench_nerf_eval.err
python - <<'PY'
import json
for f in ['r2g_bench1','r2f_bench_nerf_eval']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],4), round(d['value']), d['roofline']['frac'], d['roofline'].get('queries_per_image'))
    except Exception as e:
        print(f,'FAILED',e); print(open(f'gpurun_out/{f}.err').read()[-1200:])
PY
