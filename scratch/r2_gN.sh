#!/bin/bash
# usage: r2_gN.sh N [strong]   — the bench line on N GPUs of one box (weak scaling; optionally the strong-scaling line too)
N=$1; mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT"
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $N --steps 30 --warmup 5 $3 > gpurun_out/r2g_bench$N$1.json 2> gpurun_out/r2g_bench$N$1.err; }
run "" 29521 ""
[ "$2" = "strong" ] && run s 29522 "--scaling strong"
python - $N <<'PY'
import json, sys
N = sys.argv[1]
for f in [f'r2g_bench{N}', f'r2g_bench{N}s']:
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['scaling'], round(d['ms_per_step'], 4), round(d['value']), 'e2e', round(d['e2e']['value']), d['config']['parallelism'][:120])
    except Exception as e:
        print(f, 'FAILED', e)
        try: print(open(f'gpurun_out/{f}.err').read()[-1500:])
        except Exception: pass
PY
