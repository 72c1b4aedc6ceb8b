"""2-GPU NCCL check of the data-parallel training loop (train.reconstruction under torchrun): two ranks with
512 rays each (with an upsampling event; the sharded alpha-mask update is checked by tests/dist/dist_check.py) must reproduce the single-process run with
batch 1024.   torchrun --nproc-per-node 2 tests/dist/dist_check_train.py"""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.distributed as dist
import ffb200
from ffb200.models.FactorFields import FactorFields
from ffb200.train import reconstruction
from tests.synth_scene import sphere_scene
local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
STEPS, B = 120, 512


def make(batch):
    cfg = ffb200.load_cfg('nerf.yaml', ['model.total_params=400000', 'model.coeff_reso=16', 'training.volume_resoInit=48',
                                        'training.volume_resoFinal=64', f'training.batch_size={batch}', f'training.n_iters={STEPS}',
                                        'renderer.density_shift=-4.0'])
    cfg.dataset.aabb = [[-1., -1., -1.], [1., 1., 1.]]
    cfg.training.upsamp_list, cfg.training.update_AlphaMask_list, cfg.training.shrinking_list = [40, 10 ** 9], [10 ** 9], [10 ** 9]
    torch.manual_seed(11)
    return cfg, FactorFields(cfg, f'cuda:{local}')


rays, rgbs = sphere_scene(20000, 1)
test_rays, test_rgbs = sphere_scene(4096, 2)
T = lambda a: torch.from_numpy(a)
# single process, global batch (before the process group exists)
cfg1, m1 = make(2 * B)
np.random.seed(5); torch.manual_seed(6)
ref = reconstruction(cfg1, m1, T(rays).clone(), T(rgbs).clone(), n_iters=STEPS, test=(T(test_rays), T(test_rgbs)))
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
cfg2, m2 = make(B)
np.random.seed(5); torch.manual_seed(6)
res = reconstruction(cfg2, m2, T(rays).clone(), T(rgbs).clone(), n_iters=STEPS, test=(T(test_rays), T(test_rgbs)))
a, b = np.array(ref['psnr_train']), np.array(res['psnr_train'])
ok = abs(a[0] - b[0]) < 1e-3 and abs(a[-10:].mean() - b[-10:].mean()) < 0.1 and abs(ref['psnr_test'] - res['psnr_test']) < 0.1
print(f'rank {dist.get_rank()}: first-step PSNR single {a[0]:.4f} / sharded {b[0]:.4f}; last-10 train PSNR {a[-10:].mean():.3f} / {b[-10:].mean():.3f}; '
      f'test PSNR {ref["psnr_test"]:.3f} / {res["psnr_test"]:.3f}; alpha mask set: {m1.alphaMask is not None} / {m2.alphaMask is not None} -> {"OK" if ok else "MISMATCH"}', flush=True)
dist.barrier()
dist.destroy_process_group()
assert ok
