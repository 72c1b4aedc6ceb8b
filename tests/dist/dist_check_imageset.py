"""2-GPU NCCL check of image-set training sharded by image: each rank owns N_img / W coefficient slabs, trains on
pixels of its own images, and only the basis + MLP gradients are all-reduced.  Must equal single-process training on the
union batch (global-mean loss):  torchrun --nproc-per-node 2 tests/dist/dist_check_imageset.py"""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.distributed as dist
import ffb200
from ffb200.models.FactorFields import FactorFields
from ffb200.train import RegressStep, evaluate_field, image_set_shard
from tests import synth_scene as SS
local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank, world = dist.get_rank(), dist.get_world_size()
N_IMG, HW, B, STEPS = 6, 32, 1024, 60
SMALL = ['model.basis_dims=[8,8,8,4,4,4]', 'model.basis_resos=[8,13,18,22,27,32]', 'model.total_params=40000', 'model.with_dropout=false']
coords, imgs = SS.synth_image_set(N_IMG, HW, HW, seed=2)
coords_t, imgs_t = torch.from_numpy(coords).cuda(), torch.from_numpy(imgs).cuda()
per_img = HW * HW


def build(n_img):
    cfg = ffb200.load_cfg('image_set.yaml', SMALL)
    cfg.dataset.aabb = [[0, 0, 0], [HW, HW, n_img]]
    torch.manual_seed(3)
    return cfg, FactorFields(cfg, f'cuda:{local}')


def batch_of(step, r):
    """the B samples rank r draws at `step`: pixels of its own images only (global sample indices)"""
    i0, n = image_set_shard(N_IMG, r, world)
    g = torch.Generator().manual_seed(1000 * step + r)
    return (i0 * per_img + torch.randint(0, n * per_img, (B,), generator=g)).cuda()


# ---- reference: one process, all slabs, union batch
cfg_f, full = build(N_IMG)
init = {k: v.detach().clone() for k, v in full.state_dict().items()}
ref = RegressStep(full, full.get_optparam_groups(cfg_f.training.lr_small, cfg_f.training.lr_large), batch=B * world, x_dim=3, out_dim=3, is_train=True)
ref.world = 1                                                   # no collective on the reference side
for it in range(STEPS):
    idx = torch.cat([batch_of(it, r) for r in range(world)])
    ref.step(coords_t[idx], imgs_t[idx])
# ---- sharded: this rank's slabs only
i0, n_loc = image_set_shard(N_IMG, rank, world)
cfg_l, loc = build(n_loc)
sd = {k: (v[:, :, i0:i0 + n_loc] if k == 'coeffs.0' else v).clone() for k, v in init.items()}
loc.load_state_dict(sd)
rs = RegressStep(loc, loc.get_optparam_groups(cfg_l.training.lr_small, cfg_l.training.lr_large), batch=B, x_dim=3, out_dim=3, is_train=True,
                 local_params=list(loc.coeffs.parameters()))
shift = torch.tensor([0.0, 0.0, float(i0)], device='cuda')
for it in range(STEPS):
    idx = batch_of(it, rank)
    rs.step(coords_t[idx] - shift, imgs_t[idx])
torch.cuda.synchronize()
# ---- compare
worst = {}
for (name, pf), pl in zip(full.named_parameters(), loc.parameters()):
    a = pf.detach()[:, :, i0:i0 + n_loc] if name == 'coeffs.0' else pf.detach()
    d = (a - pl.detach()).abs()
    worst[name] = (float(d.max()), float((d > 2e-4 * max(1.0, float(a.abs().max()))).float().mean()))
own = slice(i0 * per_img, (i0 + n_loc) * per_img)
mse_ref = float(torch.mean((evaluate_field(full, coords_t[own]) - imgs_t[own]) ** 2))
mse_loc = float(torch.mean((evaluate_field(loc, coords_t[own] - shift) - imgs_t[own]) ** 2))
psnr = lambda m: -10 * np.log10(m)
ok = all(frac < 5e-3 for _, frac in worst.values()) and abs(psnr(mse_ref) - psnr(mse_loc)) < 0.05
print(f'rank {rank}: images [{i0}, {i0 + n_loc}): all-reduced arena ranges {rs._shared_ranges} of {rs.bucket.flat.numel()} floats; '
      f'PSNR on own images: single-process {psnr(mse_ref):.3f} dB, sharded {psnr(mse_loc):.3f} dB; '
      f'worst |dp| {max(v[0] for v in worst.values()):.2e}, worst out-of-tolerance fraction {max(v[1] for v in worst.values()):.2e} -> {"OK" if ok else "MISMATCH"}', flush=True)
dist.barrier()
dist.destroy_process_group()
assert ok
