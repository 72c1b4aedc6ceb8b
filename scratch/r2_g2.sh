#!/bin/bash
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; tail -3 gpurun_out/r2g_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2g_bench2.json 2> gpurun_out/r2g_bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --scaling strong > gpurun_out/r2g_bench2s.json 2> gpurun_out/r2g_bench2s.err
python - <<'PY'
import json
for f in ['r2g_bench2','r2g_bench2s']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['scaling'], round(d['ms_per_step'],4), round(d['value']), d['config']['parallelism'])
    except Exception as e:
        print(f, 'FAILED', e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
PY
