#!/bin/bash
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2g_bench2.json 2> gpurun_out/r2g_bench2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2g_bench2.json').read().strip().splitlines()[-1])
    print(d['n_gpus'], d['scaling'], round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']))
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2g_bench2.err').read()[-1500:])
PY
