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
W1, b1, W2 = m.linear_mat.backbone[0].weight, m.linear_mat.backbone[0].bias, m.linear_mat.backbone[1].weight
n = x.shape[0]
feats = _separate_fwd(m, x)['feats']
y = torch.empty(n, 32, device='cuda'); bits = torch.empty(n, 4, device='cuda', dtype=torch.int16)
gy = torch.randn(n, 32, device='cuda'); gx = torch.empty(n, 18, device='cuda')
gW1, gb1, gW2 = torch.zeros_like(W1), torch.zeros_like(b1), torch.zeros_like(W2)
P = lambda t: C.c_void_p(t.data_ptr())
for pipelined in (0, 1):
    lib.ffb_set_mlp_pipelined(pipelined)
    tf = timeit(lambda: nv.check(lib.ffb_mlp2_fwd(P(feats), P(W1), P(b1), P(W2), P(y), P(bits), C.c_int64(n), None, 18, 64, 32, nv.stream())))
    tb = timeit(lambda: nv.check(lib.ffb_mlp2_bwd(P(feats), P(gy), P(W1), P(b1), P(W2), P(bits), P(gx), P(gW1), P(gb1), P(gW2), C.c_int64(n), None, 18, 64, 32, nv.stream())))
    print('linear_mat pipelined=%d: forward %.1f us, backward %.1f us' % (pipelined, tf, tb), flush=True)
lib.ffb_set_mlp_pipelined(1)
