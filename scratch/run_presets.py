"""Drive one training step of the -vm / -CP presets at the bench shapes (for ncu captures of the specialised kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_workload as W
import ffb200
from ffb200.models.FactorFields import FactorFields
from ffb200.train import TrainStep
name = sys.argv[1]
ov, *_ = W.PRESETS[name]
import json
cfg = ffb200.load_cfg('nerf.yaml', [f'model.{k}={json.dumps(v)}' for k, v in ov.items()]); cfg.dataset.aabb = W.TNT_AABB
torch.manual_seed(0)
model = W.density_offset_(FactorFields(cfg, 'cuda:0'))
ts = TrainStep(model, model.get_optparam_groups(0.001, 0.02), batch=W.BATCH, n_samples=W.TNT_N_SAMPLES, lr_decay=0.9999, use_graph=False)
rays, target, jitter = W.tnt_rays(W.BATCH * 2, seed=100)
rays, target, jitter = (torch.from_numpy(a).cuda() for a in (rays, target, jitter))
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    s = slice((i % 2) * W.BATCH, (i % 2 + 1) * W.BATCH)
    ts.step(rays[s], target[s], jitter[s])
torch.cuda.synchronize()
print('ok', int(model.last_stats['n_valid']))
