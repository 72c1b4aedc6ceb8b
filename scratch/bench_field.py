"""Micro-benchmark of the field kernels on the bench workload: every launch configuration, CUDA events, parity vs cfg 0."""
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
g = torch.randn(n, 18, device='cuda')
def timeit(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
ref = None
for c, st in ((0, 0), (1, 0), (1, 1), (2, 1), (3, 1)):
    lib.ffb_set_tuning(b'field_fwd_cfg', c); lib.ffb_set_tuning(b'field_fwd_stage', st)
    for with_basis in (False, True):
        f = lambda: nv.check(lib.ffb_field_query_fwd_train(plan.handle, nv.ptr(x), C.c_int64(n), None, nv.ptr(feats), nv.ptr(coeff), nv.ptr(basis) if with_basis else None, nv.stream()))
        us = timeit(f)
        if ref is None: ref = feats.clone()
        print(f'fwd cfg {c} stage {st} basis={with_basis}: {us:8.1f} us  frac {n*1236/us/1e3/6553.3:.3f}  maxdiff {float((feats-ref).abs().max()):.2e}', flush=True)
lib.ffb_set_tuning(b'field_fwd_cfg', 1); lib.ffb_set_tuning(b'field_fwd_stage', 1)
cb = (coeff.clone(), basis.clone())
f(); torch.cuda.synchronize(); print('staged outputs equal direct:', bool((coeff == cb[0]).all() and (basis == cb[1]).all()), float((coeff-cb[0]).abs().max()))
grads = [torch.zeros_like(t) for t in plan.tensors]
arr = (C.c_void_p * nv.MAX_OPS)(*[gr.data_ptr() for gr in grads])
gref = None
for c, saved in ((0, False), (1, True), (2, True)):
    lib.ffb_set_tuning(b'field_bwd_cfg', c)
    f = lambda: nv.check(lib.ffb_field_query_bwd_saved(plan.handle, nv.ptr(x), C.c_int64(n), None, nv.ptr(g), None, nv.ptr(coeff) if saved else None, nv.ptr(basis) if saved else None, arr, nv.stream()))
    for gr in grads: gr.zero_()
    f(); torch.cuda.synchronize()
    cur = [gr.clone() for gr in grads]
    if gref is None: gref = cur
    err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(cur, gref))
    us = timeit(f)
    print(f'bwd cfg {c}: {us:8.1f} us  frac {n*3528/us/1e3/6553.3:.3f}  rel err vs cfg0 {err:.2e}', flush=True)
