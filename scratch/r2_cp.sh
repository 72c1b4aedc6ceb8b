#!/bin/bash
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_field.py -m gpu -x -q -k "cp or CP or preset" > gpurun_out/r2c_tests.log 2>&1; tail -3 gpurun_out/r2c_tests.log
for v in "field_lines_walk=0" "field_lines_walk=1"; do
  FFB_TUNING=$v timeout 300 python bench.py --workload nerf_cp --no-cpu-baseline --no-cuda-eager-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
  python - "$v" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2c_bench.json').read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['ms_per_step'],4), {k:v['ms_per_step'] for k,v in d['kernels'].items()})
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open('gpurun_out/r2c_bench.err').read()[-1500:])
PY
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'vm_fwd2|vm_bwd2' -s 4 -c 2 -f -o gpurun_out/r2c_prof_vm python scratch/run_presets.py nerf_vm 4 > gpurun_out/r2c_ncu_vm.log 2>&1; tail -2 gpurun_out/r2c_ncu_vm.log
