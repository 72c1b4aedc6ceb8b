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
# the CUDA-graph step's launch sequence, run eagerly: device-side counts + the sparse hand-off of the compositor's gradient
# (composite_app_fill's slot map, render_input_bwd's compact rows, mlp2p_bwd's sparse producer)
from ffb200.train import TrainStep
ts = TrainStep(m, m.get_optparam_groups(0.001, 0.02), batch=R, n_samples=S, lr_decay=0.999, use_graph=False)
tgt = torch.rand(R, 3, device='cuda')
for _ in range(2):
    ts.step(rays, tgt, torch.rand(R, device='cuda'))
torch.cuda.synchronize()
print('train step (sparse gradient hand-off:', ts._sparse_ok(), ') loss', float(ts.loss_s))
# the -CP / -vm preset kernels (field_lines.cu: TMA-staged lines + shared-memory-privatised accumulation; field_planes.cu)
nv.check(nv.lib().ffb_set_tuning(b'field_level_parallel', 1))
for ov in (['model.coeff_type=vec', 'model.basis_type=cp', 'model.freq_bands=[1.,1.,1.,1.,1.,1.]', 'model.basis_resos=[64,64,64,64,64,64]',
            'model.basis_dims=[32,32,32,32,32,32]'], ['model.coeff_type=vm', 'model.basis_type=vm']):
    c2 = ffb200.load_cfg('nerf.yaml', ['model.total_params=50000', 'model.coeff_reso=8'] + ov)
    c2.dataset.aabb = [[-1.2, -0.7, -1.0], [1.3, 0.9, 0.8]]
    m2 = FactorFields(c2, 'cuda:0')
    plan = m2._plan('coding')
    print('preset', ov[1], 'lines', nv.lib().ffb_field_lines_eligible(plan.handle), 'planes', nv.lib().ffb_field_planes_eligible(plan.handle))
    xq = torch.rand(3000, 3, device='cuda') * torch.tensor([2.5, 1.6, 1.8], device='cuda') + torch.tensor([-1.2, -0.7, -1.0], device='cuda')
    f2, _ = m2.get_coding(xq)
    g2 = torch.autograd.grad((f2 ** 2).sum(), [p for n_, p in m2.named_parameters() if n_.startswith(('coeffs', 'basises'))])
    torch.cuda.synchronize()
# the column-parallel wide-row kernels (image.yaml shapes at reduced size: 144-channel rows, nearest taps)
c3 = ffb200.load_cfg('image.yaml', ['model.total_params=90000'])
c3.dataset.aabb = [[0., 0.], [128., 128.]]
m3 = FactorFields(c3, 'cuda:0')
plan3 = m3._plan('coding')
x3 = torch.floor(torch.rand(3001, 2, device='cuda') * 128) + 0.5
f3, _ = m3.get_coding(x3)
g3 = torch.autograd.grad((f3 ** 2).sum(), [p for n_, p in m3.named_parameters() if n_.startswith(('coeffs', 'basises'))])
torch.cuda.synchronize()
print('wide rows: width', plan3.width, 'layout', nv.lib().ffb_field_saved_basis_layout(plan3.handle, 3001))
# the decoupled look-back scan (> 64 K rows)
c = torch.randint(0, 5, (70000,), device='cuda', dtype=torch.int32)
o = ops.exclusive_scan(c)
assert int(o[-1]) == int(c.sum())
print('sanitize_case done, launches', nv.launch_count())
