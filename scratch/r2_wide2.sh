#!/bin/bash
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_field.py tests/test_gpu_regress.py -m gpu -x -q -k "image or ragged or regress" > gpurun_out/r2w_tests.log 2>&1; tail -3 gpurun_out/r2w_tests.log
for v in "field_wide_pairs=0" "field_wide_pairs=1"; do
for w in image image_set; do
  FFB_TUNING=$v timeout 300 python bench.py --workload $w --no-cpu-baseline --no-cuda-eager-baseline > gpurun_out/r2w_bench_$w.json 2> gpurun_out/r2w_bench_$w.err
  python - $w $v <<'PY'
import json,sys
w=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r2w_bench_{w}.json').read().strip().splitlines()[-1])
    print(sys.argv[2], w, round(d['ms_per_step'],4), {k:v['ms_per_step'] for k,v in d['kernels'].items()}, d['roofline']['frac'])
except Exception as e:
    print(w, 'FAILED', e); print(open(f'gpurun_out/r2w_bench_{w}.err').read()[-1500:])
PY
done; done
