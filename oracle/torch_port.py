"""ORACLE / CPU BASELINE — test infrastructure, NOT product code.

Restatement of the reference's NeRF train step (grid coefficient x grid basis, bounded scene) with the SAME
third-party operators the reference calls (torch CPU: F.grid_sample, nn.functional.linear, cumprod, autograd,
Adam), so that timing it on the host cores reproduces the cost of the reference's CPU path.  `bench.py`'s NeRF legs
(cpu_baseline, `--impl reference`, the CUDA-eager bar) run the UNMODIFIED reference from the git-ignored copy under
baseline/_ref/factor-fields ("kind": "reference"); this port is their fallback when that copy is absent ("kind": "port"),
the CPU leg of the regression bench lines (the reference's regression loops live in notebooks), and the reference side of
the PSNR-parity tests.  Pinned against the golden vectors in tests/test_oracle_golden.py::test_torch_port.

Follows /root/reference/models/FactorFields.py (line numbers cited inline) and train_per_scene.py:149-171.
"""
import numpy as np
import torch
import torch.nn.functional as F


def grid_mapping_sawtooth(x, freq_bands, aabb):                          # FactorFields.py:11-22
    scale = max(aabb[1] - aabb[0])[..., None] / freq_bands
    local = (x - aabb[0])[..., None] % scale
    return (local / (scale / 2) - 1).clamp(-1., 1.)


def positional_encoding(p, freqs):                                         # :74-79
    fb = (2 ** torch.arange(freqs, device=p.device).float())
    pts = (p[..., None] * fb).reshape(p.shape[:-1] + (freqs * p.shape[-1],))
    return torch.cat([torch.sin(pts), torch.cos(pts)], dim=-1)


class TorchPort:
    """state: dict of numpy arrays with the reference's state_dict names (reference layout)."""

    def __init__(self, state, aabb, freq_bands, step_size, rcfg, lr_small=0.001, lr_large=0.02, device='cpu'):
        """device='cuda' runs the same operators as the reference's eager CUDA path (bench.py --cuda-eager-baseline);
        the pinned / default configuration is the CPU."""
        self.dev = torch.device(device)
        self.p = {k: torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(v)).float().to(self.dev)) for k, v in state.items()}
        self.aabb = torch.tensor(aabb, dtype=torch.float32, device=self.dev)
        self.freq = torch.tensor(freq_bands, dtype=torch.float32, device=self.dev)
        self.step = torch.tensor(step_size, dtype=torch.float32, device=self.dev)
        self.r = rcfg
        self.n_basis = len([k for k in state if k.startswith('basises.')])
        small = [v for k, v in self.p.items() if k.startswith(('linear_mat', 'renderModule'))]
        large = [v for k, v in self.p.items() if k.startswith(('coeffs', 'basises'))]
        self.opt = torch.optim.Adam([{'params': small, 'lr': lr_small}, {'params': large, 'lr': lr_large}], betas=(0.9, 0.99))

    def get_coding(self, x):                                               # :425-434, :467-490, :523-527
        N = x.shape[0]
        inv = 2.0 / (self.aabb[1] - self.aabb[0])
        pts = ((x - self.aabb[0]) * inv - 1).view(1, -1, 1, 1, 3)
        coeff = F.grid_sample(self.p['coeffs.0'], pts, mode='bilinear', align_corners=False, padding_mode='border').view(-1, N).t()
        xyz = grid_mapping_sawtooth(x, self.freq, self.aabb).view(1, 1, 1, -1, 3, self.freq.numel())
        bs = [F.grid_sample(self.p[f'basises.{i}'], xyz[..., i], mode='bilinear', align_corners=True).view(-1, N).T
              for i in range(self.n_basis)]
        return torch.cat(bs, dim=-1) * coeff, coeff

    def linear_mat(self, h):                                               # :144-159
        h = F.relu(F.linear(h, self.p['linear_mat.backbone.0.weight'], self.p['linear_mat.backbone.0.bias']))
        return F.linear(h, self.p['linear_mat.backbone.1.weight'])

    def render_module(self, viewdirs, features):                           # :188-203
        h = torch.cat([features, viewdirs, positional_encoding(features, self.r['fea_pe']),
                       positional_encoding(viewdirs, self.r['view_pe'])], dim=-1)
        h = F.relu(F.linear(h, self.p['renderModule.mlp.0.weight'], self.p['renderModule.mlp.0.bias']))
        h = F.relu(F.linear(h, self.p['renderModule.mlp.1.weight'], self.p['renderModule.mlp.1.bias']))
        return torch.sigmoid(F.linear(h, self.p['renderModule.mlp.2.weight']))

    def forward(self, rays, n_samples, jitter=None, white_bg=True):        # :586-602, :843-898
        o, d = rays[:, :3], rays[:, 3:6]
        vec = torch.where(d == 0, torch.full_like(d, 1e-6), d)
        t_min = torch.minimum((self.aabb[1] - o) / vec, (self.aabb[0] - o) / vec).amax(-1).clamp(min=0.05, max=1e3)
        rng = torch.arange(n_samples, device=self.dev)[None].float()
        if jitter is not None:
            rng = rng.repeat(o.shape[0], 1) + jitter[:, None]
        z = t_min[..., None] + self.step * rng
        pts = o[..., None, :] + d[..., None, :] * z[..., None]
        valid = ~((self.aabb[0] > pts) | (pts > self.aabb[1])).any(dim=-1)
        dists = torch.cat((z[:, 1:] - z[:, :-1], torch.zeros_like(z[:, :1])), dim=-1)
        viewdirs = d.view(-1, 1, 3).expand(pts.shape)
        sigma = torch.zeros(pts.shape[:-1], device=self.dev)
        rgb = torch.zeros((*pts.shape[:2], 3), device=self.dev)
        if valid.any():
            feats, _ = self.get_coding(pts[valid])
            feat = self.linear_mat(feats)
            sigma[valid] = F.softplus(feat[..., 0] + self.r['density_shift'])
        alpha = 1. - torch.exp(-sigma * dists * self.r['distance_scale'])       # :82-88
        T = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1]), 1. - alpha + 1e-10], -1), -1)
        weight = alpha * T[..., :-1]
        app = weight > self.r['rayMarch_weight_thres']
        valid_new = torch.logical_and(valid, app)
        app_c = valid_new[valid]
        if app_c.any():
            rgb[valid_new] = self.render_module(viewdirs[valid_new], feat[app_c, 1:])
        acc = torch.sum(weight, -1)
        rgb_map = torch.sum(weight[..., None] * rgb, -2)
        if white_bg:
            rgb_map = rgb_map + (1. - acc[..., None])
        self.stats = dict(n_valid=int(valid.sum()), n_app=int(valid_new.sum()))
        return rgb_map.clamp(0, 1), torch.sum(weight * z, -1).detach(), valid, weight

    def train_step(self, rays, target, n_samples, jitter):                 # train_per_scene.py:154-162
        rgb_map, _, _, _ = self.forward(rays, n_samples, jitter)
        loss = torch.mean((rgb_map - target) ** 2)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.detach())


class RegressPort:
    """The reference's regression step for grid x grid fields (scripts/2D_regression.ipynb cell 4, sdf_regression.ipynb
    cell 2, 2D_set_regression.py:120-142) restated with the torch CPU operators it calls:
    get_coding (FactorFields.py:425-434, 467-490, 523-527) -> linear_mat (:144-159) -> MSE * loss_scale -> Adam.

    state: reference-layout numpy state_dict; aabb [2, d(+1)] as `self.aabb`; in_dim 2 or 3; `images` mode when the
    coefficient tensor has one more spatial axis than in_dim (x[..., -1] selects the image slab, :287, :469-470)."""

    def __init__(self, state, aabb, freq_bands, in_dim, coef_mode='bilinear', basis_mode='bilinear', lr_small=0.001, lr_large=0.02,
                 loss_scale_decay=1.0):
        self.p = {k: torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(v)).float()) for k, v in state.items()}
        self.aabb = torch.tensor(np.asarray(aabb), dtype=torch.float32)
        self.freq = torch.tensor(np.asarray(freq_bands), dtype=torch.float32)
        self.in_dim, self.coef_mode, self.basis_mode = in_dim, coef_mode, basis_mode
        self.n_basis = len([k for k in state if k.startswith('basises.')])
        self.n_layers = len([k for k in state if k.startswith('linear_mat.backbone.') and k.endswith('.weight')])
        small = [v for k, v in self.p.items() if k.startswith('linear_mat')]
        large = [v for k, v in self.p.items() if k.startswith(('coeffs', 'basises'))]
        self.opt = torch.optim.Adam([{'params': small, 'lr': lr_small}, {'params': large, 'lr': lr_large}], betas=(0.9, 0.99))
        self.loss_scale, self.decay = 1.0, loss_scale_decay

    def get_coding(self, x):
        N, dim = x.shape
        inv = 2.0 / (self.aabb[1] - self.aabb[0])
        pts = ((x - self.aabb[0]) * inv - 1).view([1, -1] + [1] * (dim - 1) + [dim])
        coeff = F.grid_sample(self.p['coeffs.0'], pts, mode=self.coef_mode, align_corners=False, padding_mode='border').view(-1, N).t()
        xb = x[..., :self.in_dim]
        ab = self.aabb[:, :self.in_dim]
        scale = max(ab[1] - ab[0])[..., None] / self.freq                               # grid_mapping, sawtooth (:11-22)
        local = ((xb - ab[0])[..., None] % scale / (scale / 2) - 1).clamp(-1., 1.)
        xyz = local.view(1, *([1] * (self.in_dim - 1)), -1, self.in_dim, self.freq.numel())
        bs = [F.grid_sample(self.p[f'basises.{i}'], xyz[..., i], mode=self.basis_mode, align_corners=True).view(-1, N).T
              for i in range(self.n_basis)]
        return torch.cat(bs, dim=-1) * coeff, coeff

    def linear_mat(self, h, dropout_keep=None):
        if dropout_keep is not None:                                                    # F.dropout(p=0.1) with a given mask
            h = h * dropout_keep * (1.0 / 0.9)
        for l in range(self.n_layers):
            h = F.linear(h, self.p[f'linear_mat.backbone.{l}.weight'], self.p.get(f'linear_mat.backbone.{l}.bias'))
            if l != self.n_layers - 1:
                h = F.relu(h)
        return h

    def predict(self, x):
        with torch.no_grad():
            return self.linear_mat(self.get_coding(x)[0])

    def train_step(self, x, target):
        self.loss_scale *= self.decay
        y = self.linear_mat(self.get_coding(x)[0])
        loss_dist = torch.mean((y.reshape(target.shape) - target) ** 2)
        loss = loss_dist * self.loss_scale
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss_dist.detach())
