"""2-GPU NCCL check of the sharded evaluation render and the sharded alpha-lattice evaluation:
torchrun --nproc-per-node 2 tests/dist/dist_check.py  -> both must equal the single-process result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.distributed as dist
import ffb200, bench_workload as W
from ffb200.models.FactorFields import FactorFields
from ffb200.renderer import render_ray
from ffb200.train import render_sharded
local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
cfg = ffb200.load_cfg('nerf.yaml')
cfg.dataset.aabb = W.AABB
model = FactorFields(cfg, f'cuda:{local}')
model.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state(0).items()})
rays = torch.from_numpy(W.make_rays(30001, seed=5)[0])
with torch.no_grad():
    full_rgb, full_depth = render_ray(rays, model, chunk=8192, N_samples=-1, white_bg=True, is_train=False, device=model.device)
rgb, depth = render_sharded(rays, model, chunk=8192)
# chunk boundaries differ between the two runs, and small chunks take the SIMT MLP path instead of the tensor-core one:
# the maps agree to fp32 rounding, not bit for bit
err = float((rgb - full_rgb).abs().max())
ok_render = rgb.shape == full_rgb.shape and err < 1e-5 and float((depth - full_depth).abs().max()) < 1e-4
torch.manual_seed(3)
a_ref, _ = model.getDenseAlpha([40, 36, 44], times=2, sharded=False)
torch.manual_seed(3)
a_sh, _ = model.getDenseAlpha([40, 36, 44], times=2, sharded=True)
ok_alpha = torch.equal(a_ref, a_sh)
print(f'rank {dist.get_rank()}: sharded render == full render (max |d rgb| {err:.1e}): {ok_render}; sharded alpha lattice == single-process: {ok_alpha} '
      f'(occupied {float((a_ref > 0.08).float().mean()):.3f})', flush=True)
dist.barrier()
dist.destroy_process_group()
assert ok_render and ok_alpha
