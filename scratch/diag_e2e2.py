import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_workload as W
import ffb200
from ffb200.models.FactorFields import FactorFields
from ffb200.train import TrainStep
cfg = ffb200.load_cfg('nerf.yaml'); cfg.dataset.aabb = W.AABB
model = FactorFields(cfg, 'cuda:0')
model.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state(0).items()})
B, S = W.BATCH, W.N_SAMPLES
ts = TrainStep(model, model.get_optparam_groups(0.001, 0.02), batch=B, n_samples=S, lr_decay=0.9999)
r, t, j = W.make_rays(B * 4, seed=100)
rh, th, jh = (torch.from_numpy(a).pin_memory() for a in (r, t, j))
rd, td, jd = rh.cuda(), th.cuda(), jh.cuda()
T = time.perf_counter
for i in range(8): ts.step(rd[:B], td[:B], jd[:B])
torch.cuda.synchronize()
def run(host, sync, n=30):
    src = (rh, th, jh) if host else (rd, td, jd)
    marks = []
    torch.cuda.synchronize(); t0 = T()
    for i in range(n):
        sl = slice((i % 4) * B, (i % 4 + 1) * B)
        a = T()
        ts._check_params()
        ts.rays_s.copy_(src[0][sl], non_blocking=True); ts.target_s.copy_(src[1][sl], non_blocking=True); ts.jitter_s.copy_(src[2][sl], non_blocking=True)
        b = T()
        ts.graph[0].replay()
        c = T()
        if sync: v = ts.loss_s.item()
        d = T()
        marks.append((b - a, c - b, d - c))
    torch.cuda.synchronize()
    tot = (T() - t0) / n * 1e3
    m = np.array(marks).mean(0) * 1e3
    print(f'host={host} sync={sync}: {tot:.3f} ms/step; cpu: copies {m[0]:.3f} replay {m[1]:.3f} item {m[2]:.3f}', flush=True)
for host in (False, True):
    for sync in (False, True):
        run(host, sync)
# pure graph replay back to back
torch.cuda.synchronize(); t0 = T()
for i in range(30): ts.graph[0].replay()
torch.cuda.synchronize(); print('replay only', (T() - t0) / 30 * 1e3)
t0 = T()
for i in range(30):
    ts.graph[0].replay(); torch.cuda.synchronize()
print('replay + sync each', (T() - t0) / 30 * 1e3)
