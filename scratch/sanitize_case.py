"""Small drive of the shared-memory / vector-reduction kernels for compute-sanitizer (racecheck / memcheck / initcheck):
fast_fwd_kernel<LPAR=0> + fast_bwd_saved_agg_kernel (level-parallel dispatch off), mlp2_fwd/bwd, rgb_fwd/bwd, the tiled scan.
usage: compute-sanitizer --tool <tool> python scratch/sanitize_case.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ffb200
from ffb200 import native as nv, ops
from ffb200.models.FactorFields import FactorFields
from tests.golden_rays import blender_like_rays

torch.manual_seed(0)
cfg = ffb200.load_cfg('nerf.yaml', ['model.total_params=60000', 'model.coeff_reso=8'])
cfg.dataset.aabb = [[-1., -1., -1.], [1., 1., 1.]]
m = FactorFields(cfg, 'cuda:0')
with torch.no_grad():
    m.linear_mat.backbone[0].weight.mul_(10.0)
    m.linear_mat.backbone[1].weight[0].normal_(0, 2.0)
    m.linear_mat.backbone[0].weight[63].zero_()
    m.linear_mat.backbone[0].bias[63] = 1.0
    m.linear_mat.backbone[1].weight[0, 63] = 6.0
nv.check(nv.lib().ffb_set_tuning(b'field_level_parallel', 0))
R, S = 96, 120
rays = torch.from_numpy(blender_like_rays(R, 1)).cuda()
m._jitter = lambda n, tr: torch.rand(R, device='cuda')
for lazy in (False, True):
    m.lazy_counts = lazy
    rgb, depth, coeffs = m(rays, white_bg=True, is_train=True, N_samples=S)
    loss = (rgb ** 2).mean()
    grads = torch.autograd.grad(loss, list(m.parameters()), allow_unused=True)
    torch.cuda.synchronize()
    print('lazy', lazy, 'n_valid', int(m.last_stats['n_valid']), 'n_app', int(m.last_stats['n_app']), 'loss', float(loss))
# the decoupled look-back scan (> 64 K rows)
c = torch.randint(0, 5, (70000,), device='cuda', dtype=torch.int32)
o = ops.exclusive_scan(c)
assert int(o[-1]) == int(c.sum())
print('sanitize_case done, launches', nv.launch_count())
