#!/bin/bash
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_train.py tests/test_gpu_render.py -m gpu -x -q -k "sparse or train_step or psnr or render_golden or coin or scheduled" > gpurun_out/r2p_tests.log 2>&1; tail -4 gpurun_out/r2p_tests.log
for v in 0 1; do
  FFB_SPARSE_FEAT_GRAD=$v timeout 300 python bench.py --no-cpu-baseline --no-cuda-eager-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
  python - $v <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1])
    print('sparse', sys.argv[1], round(d['ms_per_step'],4), d['e2e']['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items()})
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2p_bench.err').read()[-2500:])
PY
done
