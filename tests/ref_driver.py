"""Test infrastructure: the UNMODIFIED reference (baseline/_ref/factor-fields, a copy of /root/reference staged by
__graft_entry__.build(); else /root/reference itself) driven through its own per-scene optimisation loop, restated from
train_per_scene.py:125-220 without datasets / image writers / tensorboard: rays and colours come in as tensors.  Used by the
scheduled-run parity test; never imported by the product."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available():
    return any(os.path.isdir(p) for p in (os.path.join(ROOT, 'baseline', '_ref', 'factor-fields'), '/root/reference'))


def _remove_small_objects(ar, min_size=64, connectivity=1):
    """skimage.morphology.remove_small_objects (not installed here), restated with scipy.ndimage: label with face connectivity,
    drop components smaller than min_size.  The reference calls it at FactorFields.py:774."""
    from scipy import ndimage
    labels, n = ndimage.label(ar, structure=ndimage.generate_binary_structure(ar.ndim, connectivity))
    sizes = np.bincount(labels.ravel())
    small = sizes < min_size
    small[0] = False
    out = ar.copy()
    out[small[labels]] = False
    return out


def load():
    """-> (load_cfg, FactorFields, render_ray, utils module) of the reference."""
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    from refload import load_cfg
    sys.modules['skimage.morphology'].remove_small_objects = _remove_small_objects
    sys.modules['skimage'].morphology = sys.modules['skimage.morphology']
    from models.FactorFields import FactorFields
    import renderer
    import utils
    return load_cfg, FactorFields, renderer.render_ray, utils


def reconstruction(cfg, model, render_ray, utils, allrays, allrgbs, n_iters, device='cpu', white_bg=True):
    """train_per_scene.py:125-220 (loop body verbatim in structure: same order of sampler / render / Adam / lr decay / mask
    update / shrink / ray filtering / upsample, same hard-coded `iteration >= 1500` for installing the alpha mask)."""
    import torch
    t = cfg.training
    upsamp_list, mask_list = list(t.upsamp_list), list(t.update_AlphaMask_list)
    lr_factor = t.lr_decay_target_ratio ** (1 / (t.lr_decay_iters if t.lr_decay_iters > 0 else t.n_iters))
    opt = torch.optim.Adam(model.get_optparam_groups(t.lr_small, t.lr_large), betas=(0.9, 0.99))
    reso_list = torch.linspace(t.volume_resoInit, t.volume_resoFinal, len(upsamp_list)).ceil().long().tolist()
    reso_cur = utils.N_to_reso(t.volume_resoInit ** model.in_dim, model.aabb)
    n_samples = min(cfg.renderer.max_samples, utils.cal_n_samples(reso_cur, cfg.renderer.step_ratio))
    sampler = utils.SimpleSampler(allrays.shape[0], t.batch_size)
    psnrs, reso_mask = [], None
    for it in range(n_iters):
        idx = sampler.nextids()
        rays_train, rgb_train = allrays[idx], allrgbs[idx].to(device)
        rgb_map, _, _ = render_ray(rays_train, model, chunk=t.batch_size, N_samples=n_samples, white_bg=white_bg, ndc_ray=False,
                                   device=device, is_train=True)
        loss = torch.mean((rgb_map - rgb_train) ** 2)
        opt.zero_grad()
        loss.backward()
        opt.step()
        psnrs.append(-10.0 * np.log(loss.detach().item()) / np.log(10.0))
        for g in opt.param_groups:
            g['lr'] = g['lr'] * lr_factor
        if it in mask_list or it in t.shrinking_list:
            if reso_list[0] < 256:
                reso_mask = utils.N_to_reso(reso_list[0] ** model.in_dim, model.aabb)
            new_aabb = model.updateAlphaMask(tuple(reso_mask), is_update_alphaMask=it >= 1500)
            if it in t.shrinking_list:
                model.shrink(new_aabb)
                opt = torch.optim.Adam(model.get_optparam_groups(t.lr_small, t.lr_large), betas=(0.9, 0.99))
            if not cfg.dataset.ndc_ray and it == mask_list[0] and not cfg.dataset.is_unbound:
                allrays, allrgbs = model.filtering_rays(allrays, allrgbs)
                sampler = utils.SimpleSampler(allrgbs.shape[0], t.batch_size)
        if it in upsamp_list:
            n_voxels = reso_list.pop(0)
            reso_cur = utils.N_to_reso(n_voxels ** model.in_dim, model.aabb)
            n_samples = min(cfg.renderer.max_samples, utils.cal_n_samples(reso_cur, cfg.renderer.step_ratio))
            model.upsample_volume_grid(reso_cur)
            opt = torch.optim.Adam(model.get_optparam_groups(t.lr_small, t.lr_large), betas=(0.9, 0.99))
    return dict(psnr_train=psnrs, n_rays=int(allrays.shape[0]), n_samples=n_samples)
