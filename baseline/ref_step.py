"""Reference arm of bench.py — NOT product source and NOT a port.

Runs the UNMODIFIED reference checkout (baseline/_ref/factor-fields, a git-ignored copy of /root/reference made by
`__graft_entry__.build()`; it travels to the GPU box with the working tree) through its own public API:

    models.FactorFields.FactorFields(cfg, device)  +  renderer.render_ray(..., is_train=True)
    loss = mean((rgb_map - rgb_train)**2);  optimizer.zero_grad();  loss.backward();  optimizer.step()
    loss.detach().item();  lr *= lr_factor                                (train_per_scene.py:149-171)

on the bench workload of bench_workload.py (nerf.yaml shapes, seeded synthetic state and rays).  The only things
supplied from outside are what the training script reads from a dataset: the rays, the target colours, and (for
reproducibility) the seed of the torch CPU generator the reference draws its per-ray jitter from.
`device='cpu'` is the reference's CPU path (cpu_baseline / --impl reference, kind "reference");
`device='cuda'` is the reference's eager CUDA path on the same B200 (the kernel-level bar of SURVEY 8(d)).
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return any(os.path.isdir(p) for p in (os.environ.get('FF_REF') or '', os.path.join(_HERE, '_ref', 'factor-fields'), '/root/reference'))


class RefStep:
    def __init__(self, state, aabb, device, n_samples, overrides=None, lr_small=0.001, lr_large=0.02, n_iters=30000, seed=20211202,
                 cfgname='nerf.yaml', dataset_overrides=None):
        if _HERE not in sys.path:
            sys.path.insert(0, _HERE)
        import numpy as np
        import torch
        from refload import load_cfg          # puts the reference checkout on sys.path, stubs the non-hot-path imports
        from models.FactorFields import FactorFields
        import renderer as R
        self.torch, self.R = torch, R
        torch.manual_seed(seed)
        np.random.seed(seed)
        cfg = load_cfg(cfgname)
        cfg.dataset.aabb = aabb
        for k, v in (overrides or {}).items():
            cfg.model[k] = v
        for k, v in (dataset_overrides or {}).items():
            cfg.dataset[k] = v
        self.cfg, self.dev, self.S = cfg, device, int(n_samples)
        self.model = FactorFields(cfg, device)
        if state is not None:
            sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in state.items()}
            self.model.load_state_dict(sd)
        self.model.to(device)
        self.opt = torch.optim.Adam(self.model.get_optparam_groups(lr_small, lr_large), betas=(0.9, 0.99))
        self.lr_factor = 0.1 ** (1.0 / n_iters)
        self.stats = {}

    def train_step(self, rays_host, target):
        """rays_host: host tensor [n, 6] (the reference keeps rays on the host, train_per_scene.py:146,151-152);
        target [n, 3] on the device, like `allrgbs[ray_idx].to(device)`."""
        torch = self.torch
        n = rays_host.shape[0]
        rgb_map, depth_map, coeffs = self.R.render_ray(rays_host, self.model, chunk=n, N_samples=self.S, white_bg=True, ndc_ray=False,
                                                       device=self.dev, is_train=True)
        loss = torch.mean((rgb_map - target) ** 2)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        loss = loss.detach().item()
        for g in self.opt.param_groups:
            g['lr'] = g['lr'] * self.lr_factor
        self.stats = {'n_valid': int(coeffs.shape[0]), 'loss': float(loss)}
        return loss
