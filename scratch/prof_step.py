import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_workload as W
import ffb200
from ffb200.models.FactorFields import FactorFields
from ffb200.train import TrainStep
from torch.profiler import profile, ProfilerActivity
cfg = ffb200.load_cfg('nerf.yaml'); cfg.dataset.aabb = W.AABB
model = FactorFields(cfg, 'cuda:0')
model.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state(0).items()})
B, S = W.BATCH, W.N_SAMPLES
ts = TrainStep(model, model.get_optparam_groups(0.001, 0.02), batch=B, n_samples=S, lr_decay=0.9999, use_graph=False)
rays, target, jitter = W.make_rays(B * 4, seed=100)
rays, target, jitter = (torch.from_numpy(a).cuda() for a in (rays, target, jitter))
for i in range(6): ts.step(rays[(i%4)*B:(i%4+1)*B], target[(i%4)*B:(i%4+1)*B], jitter[(i%4)*B:(i%4+1)*B])
torch.cuda.synchronize()
N = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(N): ts.step(rays[(i%4)*B:(i%4+1)*B], target[(i%4)*B:(i%4+1)*B], jitter[(i%4)*B:(i%4+1)*B])
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.OrderedDict()
for e in ev:
    a = agg.setdefault(e.name[:100], [0, 0.0]); a[0] += 1; a[1] += e.time_range.elapsed_us()
tot = sum(a[1] for a in agg.values())
print(f'n_valid {int(model.last_stats["n_valid"])} n_app {int(model.last_stats["n_app"])}; total kernel time per step {tot/N:.1f} us')
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:28]:
    print(f'{a[0]/N:5.1f} x {a[1]/a[0]:8.1f} us = {a[1]/N:8.1f} us/step  {k}')
