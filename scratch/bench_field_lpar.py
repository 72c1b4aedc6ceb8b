"""A/B: one thread per query vs one thread per (query, level) for the field forward at the nerf.yaml batch (0.99 M queries)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_workload as W
import ffb200
from ffb200 import native as nv, ops
from ffb200.models.FactorFields import FactorFields
cfg = ffb200.load_cfg('nerf.yaml'); cfg.dataset.aabb = W.AABB
m = FactorFields(cfg, 'cuda:0')
m.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state(0).items()})
rays, target, jitter = W.make_rays(W.BATCH, seed=100)
samp = ops.sample_compact(m._sampler_desc(W.N_SAMPLES, False), torch.from_numpy(rays).cuda(), torch.from_numpy(jitter).cuda())
x = samp['xyz']; n = x.shape[0]
plan = m._plan('coding'); lib = nv.lib()
feats, coeff = (torch.empty(n, 18, device='cuda') for _ in range(2))
basis = torch.empty((n + 31) // 32 * 32, 18, device='cuda')      # blocked by 32 rows
def timeit(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
ref = None
for lpar in (0, 1):
    for cfgk in (1, 2, 3):
        lib.ffb_set_tuning(b'field_fwd_lpar_all', lpar); lib.ffb_set_tuning(b'field_fwd_cfg', cfgk)
        f = lambda: nv.check(lib.ffb_field_query_fwd_train(plan.handle, nv.ptr(x), C.c_int64(n), None, nv.ptr(feats), nv.ptr(coeff), nv.ptr(basis), nv.stream()))
        us = timeit(f)
        if ref is None: ref = feats.clone()
        print(f'level-parallel {lpar} cfg {cfgk}: {us:8.1f} us  frac {n*1236/us/1e3/6553.3:.3f}  maxdiff {float((feats-ref).abs().max()):.2e}', flush=True)
