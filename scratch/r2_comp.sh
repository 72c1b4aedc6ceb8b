#!/bin/bash
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_train.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r2k_tests.log 2>&1; tail -3 gpurun_out/r2k_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-cuda-eager-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
python - <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1])
    print(round(d['ms_per_step'],4), d['e2e']['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items()})
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2k_bench.err').read()[-1500:])
PY
