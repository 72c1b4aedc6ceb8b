"""Fused field + linear_mat kernels vs the separate kernels at the bench workload (0.99 M ray-ordered queries): CUDA-event times."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_workload as W
import ffb200
from ffb200 import native as nv, ops
from ffb200.models.FactorFields import FactorFields
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.test_gpu_field_mlp import _fused_fwd, _separate_fwd

cfg = ffb200.load_cfg('nerf.yaml'); cfg.dataset.aabb = W.AABB
m = FactorFields(cfg, 'cuda:0')
m.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state(0).items()})
rays, _, jitter = W.make_rays(W.BATCH, seed=100)
samp = ops.sample_compact(m._sampler_desc(W.N_SAMPLES, False), torch.from_numpy(rays).cuda(), torch.from_numpy(jitter).cuda())
x = samp['xyz']
print('queries', x.shape[0])


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print('separate fwd (field + mlp2)  %.1f us' % timeit(lambda: _separate_fwd(m, x)))
print('fused fwd (training outputs) %.1f us' % timeit(lambda: _fused_fwd(m, x, want_rows=False)))
print('fused fwd (+ row-major rows) %.1f us' % timeit(lambda: _fused_fwd(m, x, want_rows=True)))

lib = nv.lib()
for nslot in (8, 6, 4):
    for dbg in (0, 2, 1):
        lib.ffb_field_mlp_tuning(dbg, nslot)
        print('nslot %d (gather warps %d) debug %d (1: no gather math, 2: no epilogue work): %.1f us' % (nslot, min(27, 4 * nslot), dbg, timeit(lambda: _fused_fwd(m, x, want_rows=False))), flush=True)
lib.ffb_field_mlp_tuning(0, 8)
for cfgk in (1, 3, 0):
    lib.ffb_set_tuning(b'field_fwd_cfg', cfgk)
    print('old field fwd cfg %d + mlp2: %.1f us' % (cfgk, timeit(lambda: _separate_fwd(m, x))), flush=True)
