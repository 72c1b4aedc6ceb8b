"""B200-native host module with the call surface of the reference's `models/FactorFields.py`.

Same class / method / attribute names, config keys, parameter names and shapes (state_dict compatible)
as the reference; everything numeric below the Python surface runs in libffb200.so (sm_100a CUDA).
Factor tensors keep the reference's logical shape [1,C,(D,)H,W] but are stored channels-last, which is
what the kernels read.  There is no CPU path: a non-CUDA device raises.

Reference line numbers cited as FactorFields.py:N refer to /root/reference/models/FactorFields.py.
"""
import ctypes as C
import math
import time

import numpy as np
import torch
import torch.nn

from .. import native as nv
from .. import ops
from ..utils import N_to_reso, N_to_vm_reso, remove_small_objects


def _channels_last(t):
    if t.dim() == 5:
        return t.contiguous(memory_format=torch.channels_last_3d)
    if t.dim() == 4:
        return t.contiguous(memory_format=torch.channels_last)
    return t.contiguous()


def _factor_param(t):
    return torch.nn.Parameter(_channels_last(t))


# ---- free functions of the reference module ---------------------------------------------------------
def grid_mapping(positions, freq_bands, aabb, basis_mapping='sawtooth'):
    """FactorFields.py:11-33"""
    return ops.grid_mapping(positions, freq_bands, aabb, basis_mapping)


def dct_dict(n_atoms_fre, size, n_selete, dim=2):
    """FactorFields.py:36-71 — separable DCT dictionary used to initialise grid bases: rows cos(k*pi*i/p), mean
    removed for k>0, Kronecker product over `dim` axes, n_selete rows picked evenly, each L2-normalised."""
    p = n_atoms_fre
    i = np.arange(size)
    rows = np.stack([np.cos(i * k * math.pi / p) for k in range(p)])
    rows[1:] -= rows[1:].mean(axis=1, keepdims=True)
    atoms = np.kron(rows, rows)
    if dim == 3:
        atoms = np.kron(atoms, rows)
    if n_selete < atoms.shape[0]:
        pick = [chunk[0] for chunk in np.array_split(np.arange(atoms.shape[0]), n_selete)]
        atoms = atoms[pick]
    for r in range(atoms.shape[0]):
        atoms[r] /= (np.linalg.norm(atoms[r]) or 1)
    return torch.FloatTensor(atoms)


@nv.on_device_of(0)
def positional_encoding(positions, freqs):
    """FactorFields.py:74-79 (device kernel: ffb_pe_concat_fwd without the identity part)."""
    shp = positions.shape
    x = positions.reshape(-1, shp[-1]).contiguous().float()
    D = shp[-1]
    out = torch.empty((x.shape[0], D + 2 * D * freqs), device=x.device)
    if x.shape[0] > 0:
        nv.check(nv.lib().ffb_pe_concat_fwd(nv.ptr(x), nv.ptr(out), C.c_int64(x.shape[0]), None, D, freqs, nv.stream()))
    return out[:, D:].reshape(*shp[:-1], 2 * D * freqs)


@torch.no_grad()
@nv.on_device_of(0)
def raw2alpha(sigma, dist):
    """FactorFields.py:82-88 on dense [rays, samples] inputs -> (alpha, weights, T[..., -1:]); no autograd (forward()
    differentiates through ops.RenderComposite instead).  Runs the composite kernel with every sample marked valid."""
    R, S = sigma.shape
    offsets = (torch.arange(R + 1, device=sigma.device, dtype=torch.int32) * S).contiguous()
    cdesc = ops.make_composite_desc(0.0, 'relu', 1.0, 1e30, True)   # relu(sigma + 0) == sigma for sigma >= 0
    s = sigma.reshape(-1).contiguous().float()
    d = dist.reshape(-1).contiguous().float()
    sg, tr, w = torch.empty_like(s), torch.empty_like(s), torch.empty_like(s)
    cnt = torch.empty(R, device=s.device, dtype=torch.int32)
    nv.check(nv.lib().ffb_composite_weights(C.byref(cdesc), nv.ptr(s), 1, nv.ptr(d), nv.i32p(offsets), C.c_int64(R), nv.ptr(sg),
                                            nv.ptr(tr), nv.ptr(w), nv.i32p(cnt), nv.stream()))
    w = w.view(R, S)
    tr = tr.view(R, S)
    alpha = 1. - torch.exp(-sigma * dist)
    return alpha, w, tr[:, -1:] * (1. - alpha[:, -1:] + 1e-10)


class AlphaGridMask(torch.nn.Module):
    """FactorFields.py:91-110"""

    def __init__(self, device, aabb, alpha_volume):
        super().__init__()
        self.device = device
        self.aabb = aabb.to(self.device)
        self.aabbSize = self.aabb[1] - self.aabb[0]
        self.invgridSize = 1.0 / self.aabbSize * 2
        self.alpha_volume = alpha_volume.view(1, 1, *alpha_volume.shape[-3:]).to(self.device)
        self.gridSize = torch.LongTensor([alpha_volume.shape[-1], alpha_volume.shape[-2], alpha_volume.shape[-3]]).to(self.device)
        # kernel-side copy: the volume only ever holds 0/1 (FactorFields.py:772-776), one byte per voxel
        self.volume_u8 = (self.alpha_volume[0, 0] > 0.5).to(torch.uint8).contiguous()
        if not bool(((self.alpha_volume == 0) | (self.alpha_volume == 1)).all()):
            raise RuntimeError('AlphaGridMask expects a 0/1 volume')

    @nv.on_device_of(1)
    def sample_alpha(self, xyz_sampled):
        xyz = xyz_sampled.reshape(-1, 3).contiguous().float()
        desc = ops.make_sampler_desc(self.aabb, 0.0, 1, alpha=self)
        out = torch.empty(xyz.shape[0], device=xyz.device)
        if xyz.shape[0] > 0:
            nv.check(nv.lib().ffb_alpha_sample(C.byref(desc), nv.ptr(xyz), C.c_int64(xyz.shape[0]), nv.ptr(out), nv.stream()))
        return out

    def normalize_coord(self, xyz_sampled):
        return (xyz_sampled - self.aabb[0]) * self.invgridSize - 1


class MLPMixer(torch.nn.Module):
    """FactorFields.py:113-159"""

    def __init__(self, in_dim, out_dim=16, num_layers=2, hidden_dim=64, pe=0, with_dropout=False):
        super().__init__()
        self.with_dropout = with_dropout
        self.in_dim = in_dim + 2 * in_dim * pe
        self.num_layers, self.hidden_dim, self.pe = num_layers, hidden_dim, pe
        layers = []
        for l in range(num_layers):
            last = l == num_layers - 1
            layers.append(torch.nn.Linear(self.in_dim if l == 0 else hidden_dim, out_dim if last else hidden_dim, bias=not last))
        self.backbone = torch.nn.ModuleList(layers)

    def _flat(self):
        params, has_bias = [], []
        for lin in self.backbone:
            params.append(lin.weight)
            has_bias.append(lin.bias is not None)
            if lin.bias is not None:
                params.append(lin.bias)
        return params, tuple(has_bias)

    def forward(self, x, is_train=False, n_dev=None):
        lead = x.shape[:-1]
        h = x.reshape(-1, x.shape[-1])
        params, has_bias = self._flat()
        if self.with_dropout and is_train:
            # F.dropout(p=0.1) sits between the PE concat and the first layer (FactorFields.py:147-151)
            if self.pe > 0:
                h = torch.cat([h, positional_encoding(h, self.pe)], dim=-1)
            keep = (torch.rand_like(h) >= 0.1).to(h.dtype)
            h = h * keep * (1.0 / 0.9)
            out = ops.MLPFunction.apply(h, 0, has_bias, n_dev, *params)
        else:
            out = ops.MLPFunction.apply(h, self.pe, has_bias, n_dev, *params)
        return out.reshape(*lead, out.shape[-1])


class MLPRender_Fea(torch.nn.Module):
    """FactorFields.py:162-203"""

    def __init__(self, inChanel, num_layers=3, hidden_dim=64, viewpe=6, feape=2):
        super().__init__()
        self.in_mlpC = 3 + inChanel + 2 * viewpe * 3 + 2 * feape * inChanel
        self.num_layers, self.viewpe, self.feape = num_layers, viewpe, feape
        layers = []
        for l in range(num_layers):
            last = l == num_layers - 1
            layers.append(torch.nn.Linear(self.in_mlpC if l == 0 else hidden_dim, 3 if last else hidden_dim, bias=not last))
        self.mlp = torch.nn.ModuleList(layers)

    def _flat(self):
        params, has_bias = [], []
        for lin in self.mlp:
            params.append(lin.weight)
            has_bias.append(lin.bias is not None)
            if lin.bias is not None:
                params.append(lin.bias)
        return params, tuple(has_bias)

    def forward(self, viewdirs, features):
        params, has_bias = self._flat()
        return ops.RenderMLP.apply(viewdirs.contiguous().float(), features.contiguous().float(), self.viewpe, self.feape, has_bias,
                                   *params)


def field_shapes(cfg, aabb, device='cpu'):
    """Pure host logic of FactorFields.setup_params (FactorFields.py:262-307): every grid shape, the frequency bands
    and the parameter budget split, as a dict of the attributes the reference sets on `self`."""
    m, mode = cfg.model, cfg.defaults.mode
    coeff_type, basis_type = m.coeff_type, m.basis_type
    out = {}
    multi = mode in ('images', 'reconstructions')
    d = len(aabb[0]) - 1 if multi else len(aabb[0])
    out['in_dim'] = d
    box = torch.FloatTensor(aabb)[:, :d].to(device)
    out['aabb'] = box
    out['basis_dims'] = m.basis_dims
    dims = np.array(m.basis_dims)
    width = sum(m.basis_dims)
    factorised_coeff = any(k in coeff_type for k in ('vec', 'cp', 'vm'))
    if 'reconstruction' not in mode:
        if 'image' in mode:    # NB: substring test, true for 'images' too
            basis_reso = m.basis_resos
        else:
            basis_reso = np.round(np.array(m.basis_resos) * (min(aabb[1][:d]) + 1) / 1024.0).astype('int').tolist()
        T_basis = m.T_basis if m.T_basis > 0 else sum(np.power(np.array(basis_reso), d) * dims)
        T_coeff = m.T_coeff if m.T_coeff > 0 else m.total_params - T_basis
        if not T_coeff > 0:
            T_coeff = 8 ** d * width
        if mode == 'image':
            freq_bands = max(aabb[1][:d]) / torch.FloatTensor(basis_reso).to(device)
        else:
            freq_bands = torch.FloatTensor(m.freq_bands).to(device)
        coeff_reso = N_to_reso(T_coeff // width, box[:, :d])[::-1]   # D, H, W
        if mode == 'sdf':
            freq_bands *= 0.5
        elif mode == 'images':
            coeff_reso = [aabb[1][-1]] + coeff_reso
            out['aabb'] = torch.FloatTensor(aabb).to(device)
        if factorised_coeff:
            coeff_reso = aabb[1]
        n_scene = 1
    else:
        coeff_reso = N_to_reso(m.coeff_reso ** d, box[:, :d])[::-1]
        T_coeff = width * np.prod(coeff_reso)
        T_basis = m.total_params - T_coeff
        ratio = T_basis / sum(np.power(np.array(m.basis_resos), d) * dims)
        ratio = np.power(ratio, 1.0 / d)
        line_basis = 'vec' in basis_type or 'cp' in basis_type
        basis_reso = m.basis_resos if line_basis else np.round(np.array(m.basis_resos) * ratio).astype('int').tolist()
        freq_bands = torch.FloatTensor(m.freq_bands).to(device)
        if not (mode == 'reconstructions' or 'x' in basis_type or line_basis):
            freq_bands = freq_bands * (cfg.dataset.scene_reso / float(max(basis_reso)) / max(m.freq_bands))
        n_scene = int(aabb[1][-1]) if mode == 'reconstructions' else 1
    out.update(basis_reso=basis_reso, T_basis=T_basis, T_coeff=T_coeff, freq_bands=freq_bands, coeff_reso=coeff_reso,
               n_scene=n_scene)
    return out


class FactorFields(torch.nn.Module):
    """FactorFields.py:206-898"""

    def __init__(self, cfg, device):
        super().__init__()
        self.cfg = cfg
        self.device = device
        if not str(device).startswith('cuda'):
            raise RuntimeError('factor-fields_b200 runs on CUDA devices only (no CPU fallback); got device=%r' % (device,))
        nv.lib()  # fail loudly right here if the CUDA library is missing
        self.matMode = [[0, 1], [0, 2], [1, 2]]
        self.vecMode = [2, 1, 0]
        self.n_scene, self.scene_idx = 1, 0
        self.alphaMask = None
        self._plans = {}
        self.coeff_type, self.basis_type = cfg.model.coeff_type, cfg.model.basis_type

        self.setup_params(self.cfg.dataset.aabb)
        if self.cfg.model.coeff_type != 'none':
            self.coeffs = self.init_coef()
        if self.cfg.model.basis_type != 'none':
            self.basises = self.init_basis()

        n_levels, width = len(cfg.model.basis_dims), sum(cfg.model.basis_dims)
        if 'vm' in self.coeff_type:
            mlp_in = width * 3
        elif 'x' in self.cfg.model.basis_type:   # NB: also true for 'fix-grid' 
            mlp_in = n_levels * self.in_dim * (2 if self.cfg.model.basis_mapping == 'trigonometric' else 1)
        else:
            mlp_in = width
        out_dim = cfg.model.out_dim
        self.linear_mat = MLPMixer(mlp_in, out_dim, num_layers=cfg.model.num_layers, hidden_dim=cfg.model.hidden_dim,
                                   with_dropout=cfg.model.with_dropout).to(device)

        if 'reconstruction' in cfg.defaults.mode:
            r = cfg.renderer
            self.renderModule = MLPRender_Fea(inChanel=out_dim - 1, num_layers=r.num_layers, hidden_dim=r.hidden_dim,
                                              viewpe=r.view_pe, feape=r.fea_pe).to(device)
            self.is_unbound = self.cfg.dataset.is_unbound
            if self.is_unbound:
                self.bg_len = 0.2
                self.inward_aabb = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]]).to(device)
                self.aabb = self.inward_aabb * (1 + self.bg_len)
            else:
                self.inward_aabb = self.aabb
            self.cur_volumeSize = N_to_reso(cfg.training.volume_resoInit ** self.in_dim, self.aabb)
            self.update_renderParams(self.cur_volumeSize)

        print('=====> total parameters: ', self.n_parameters())

    # ---- shape logic (FactorFields.py:262-307) ------------------------------------------------------
    def setup_params(self, aabb):
        for k, v in field_shapes(self.cfg, aabb, self.device).items():
            setattr(self, k, v)
        self._plans = {}

    def init_coef(self):
        """FactorFields.py:311-332"""
        mode, init = self.cfg.defaults.mode, self.cfg.model.coef_init
        width = sum(self.basis_dims)
        n_scene = self.n_scene if mode in ('reconstructions', 'images') else 1
        ct = self.coeff_type
        if 'hash' in ct:
            raise NotImplementedError("coeff_type 'hash' needs tiny-cuda-nn, which north_star excludes")
        if 'grid' in ct:
            shape = (1, width, *[int(r) for r in self.coeff_reso])
            return torch.nn.ParameterList([_factor_param(init * torch.ones(shape, device=self.device)) for _ in range(n_scene)])
        if 'cp' in ct or 'vm' in ct:
            return torch.nn.ParameterList([_factor_param(init * torch.ones((1, width, int(max(256, r)), n_scene), device=self.device))
                                           for r in self.coeff_reso])
        if 'vec' in ct:
            return torch.nn.ParameterList([_factor_param(init * torch.ones((1, width, int(max(256, max(self.coeff_reso))), n_scene),
                                                                            device=self.device))])
        if 'mlp' in ct:
            return torch.nn.ModuleList([MLPMixer(self.in_dim, width, num_layers=2, hidden_dim=64, pe=4).to(self.device)
                                        for _ in range(n_scene)])
        raise ValueError('unknown coeff_type %r' % ct)

    def init_basis(self):
        """FactorFields.py:334-423"""
        bt, d = self.basis_type, self.in_dim
        if 'hash' in bt:
            raise NotImplementedError("basis_type 'hash' needs tiny-cuda-nn, which north_star excludes")
        if 'mlp' in bt:
            return torch.nn.ModuleList([MLPMixer(d, c, num_layers=2, hidden_dim=64, pe=4).to(self.device) for c in self.basis_dims])
        out = []
        for c, reso in zip(self.basis_dims, self.basis_reso):
            if 'grid' in bt:
                atoms = dct_dict(int(np.power(c, 1. / d) + 1), reso, n_selete=c, dim=d)
                out.append(_factor_param(atoms.reshape([1, c] + [reso] * d).to(self.device)))
            elif 'vm' in bt:
                plane = N_to_vm_reso(reso ** d, self.aabb[:, :d])
                for a0, a1 in self.matMode:
                    out.append(_factor_param(0.1 * torch.randn((1, c, plane[a1], plane[a0]), device=self.device)))
            elif 'cp' in bt:
                for _ in range(d - 1):
                    out.append(_factor_param(0.1 * torch.randn((1, c, max(reso, 128), 1), device=self.device)))
            elif 'x' in bt:
                continue
        return torch.nn.ParameterList(out)

    # ---- gather plans -------------------------------------------------------------------------------
    def _coeff_is_mlp(self):
        return self.cfg.model.coeff_type != 'none' and not any(k in self.coeff_type for k in ('grid', 'vec', 'cp', 'vm')) \
            and 'mlp' in self.coeff_type

    def _basis_is_mlp(self):
        return self.cfg.model.basis_type != 'none' and 'mlp' in self.basis_type

    def _add_coeff_ops(self, pb):
        """get_coeff (FactorFields.py:425-465) as gather terms; -> width"""
        ct, d = self.coeff_type, self.in_dim
        nearest = self.cfg.model.coef_mode == 'nearest'
        col_const = (self.scene_idx + 0.5) / self.n_scene * 2 - 1
        kw = dict(space=0, align=False, border=True, nearest=nearest)
        if 'grid' in ct:
            t = self.coeffs[self.scene_idx]
            nd = t.dim() - 2
            pb.term('c', [pb.op(t, list(range(nd)), **kw)], 0)
            return t.shape[1]
        if 'vec' in ct:
            t = self.coeffs[0]
            pb.term('c', [pb.op(t, [-1, 0], cst=(col_const, 0, 0), **kw)], 0)
            return t.shape[1]
        if 'cp' in ct:
            pb.term('c', [pb.op(self.coeffs[i], [-1, i], cst=(col_const, 0, 0), **kw) for i in range(d)], 0)
            return self.coeffs[0].shape[1]
        if 'vm' in ct:
            col = 0
            for i in range(d):
                t = self.coeffs[i]
                pb.term('c', [pb.op(t, [-1, self.vecMode[i]], cst=(col_const, 0, 0), **kw)], col)
                col += t.shape[1]
            return col
        raise ValueError('unknown coeff_type %r' % ct)

    def _add_basis_ops(self, pb):
        """get_basis (FactorFields.py:467-516) as gather terms; -> (width, is_x, perm)"""
        bt, d = self.basis_type, self.in_dim
        F = len(self.freq_bands)
        nearest = self.cfg.model.basis_mode == 'nearest'
        col = 0
        if 'grid' in bt:
            for i in range(F):
                t = self.basises[i]
                pb.term('b', [pb.op(t, list(range(d)), space=1, level=i, align=True, border=False, nearest=nearest)], col)
                col += t.shape[1]
            return col, False, None
        if 'vm' in bt:
            for i in range(F):
                for m in range(d):
                    t = self.basises[i * d + m]
                    pb.term('b', [pb.op(t, list(self.matMode[m]), space=1, level=i, align=True, border=False)], col)
                    col += t.shape[1]
            per = col // F
            q = np.arange(col)
            return col, False, (q % per) * F + q // per      # :514-515  view(N,F,-1).permute(0,2,1)
        if 'cp' in bt:
            for i in range(F):
                ops_ = [pb.op(self.basises[i * (d - 1) + a], [-1, a + 1], space=1, level=i, align=True, border=False)
                        for a in range(d - 1)]
                pb.term('b', ops_, col)
                col += self.basises[i * (d - 1)].shape[1]
            return col, False, None
        if 'x' in bt:
            return F * d * (2 if self.cfg.model.basis_mapping == 'trigonometric' else 1), True, None
        raise ValueError('unknown basis_type %r' % bt)

    def _plan(self, which):
        """which in {'coding', 'coeff', 'basis'}; cached until a factor tensor is re-allocated."""
        key = (which, self.scene_idx)
        plan = self._plans.get(key)
        if plan is not None and not plan.stale():
            return plan
        images = self.cfg.defaults.mode == 'images'
        xdim = self.in_dim + 1 if images else self.in_dim
        pb = ops.PlanBuilder(xdim, self.in_dim, self.aabb, self.cfg.model.basis_mapping, self.freq_bands)
        cw = bw = 0
        is_x, perm = False, None
        if which in ('coding', 'coeff') and self.cfg.model.coeff_type != 'none':
            cw = self._add_coeff_ops(pb)
        if which in ('coding', 'basis') and self.cfg.model.basis_type != 'none':
            bw, is_x, perm = self._add_basis_ops(pb)
        plan = pb.finish(cw, bw, is_x, perm, self.device)
        self._plans[key] = plan
        return plan

    # ---- field queries ------------------------------------------------------------------------------
    def get_coeff(self, xyz_sampled):
        """FactorFields.py:425-465"""
        if self._coeff_is_mlp():
            return self.coeffs[self.scene_idx](self.normalize_coord(xyz_sampled))
        plan = self._plan('coeff')
        return ops.FieldQuery.apply(plan, xyz_sampled, None, *plan.tensors)[1]

    def get_basis(self, x):
        """FactorFields.py:467-516"""
        if self._basis_is_mlp():
            if self.cfg.defaults.mode == 'images':
                x = x[..., :-1]
            F = len(self.freq_bands)
            xyz = grid_mapping(x, self.freq_bands, self.aabb[:, :self.in_dim], self.cfg.model.basis_mapping).reshape(-1, self.in_dim, F)
            return torch.cat([self.basises[i](xyz[..., i].reshape(-1, self.in_dim)) for i in range(F)], dim=-1)
        plan = self._plan('basis')
        return ops.FieldQuery.apply(plan, x, None, *plan.tensors)[0]

    @torch.no_grad()
    def normalize_basis(self):
        """FactorFields.py:518-521"""
        for basis in self.basises:
            basis.data = _channels_last(basis.data / torch.norm(basis.data, dim=(2, 3), keepdim=True))

    def get_coding(self, x, n_dev=None):
        """FactorFields.py:523-533 — one fused kernel when both factors are tensors.  n_dev: optional device-side row
        count (x is then a capacity-sized buffer; used by forward() in lazy mode)."""
        has_c, has_b = self.cfg.model.coeff_type != 'none', self.cfg.model.basis_type != 'none'
        if (has_c and self._coeff_is_mlp()) or (has_b and self._basis_is_mlp()):
            if has_c and has_b:
                coeff = self.get_coeff(x)
                return self.get_basis(x) * coeff, coeff
            only = self.get_coeff(x) if has_c else self.get_basis(x)
            return only, only
        plan = self._plan('coding')
        feats, coeff = ops.FieldQuery.apply(plan, x, n_dev, *plan.tensors)
        if has_c and has_b:
            return feats, coeff
        return feats, feats

    def n_parameters(self):
        total = sum(p.numel() for p in self.parameters())
        if 'fix' in self.cfg.model.basis_type:
            total -= self.T_basis
        return total

    def get_optparam_groups(self, lr_small=0.001, lr_large=0.02):
        """FactorFields.py:541-554"""
        t, groups = self.cfg.training, []
        if t.linear_mat:
            groups.append({'params': self.linear_mat.parameters(), 'lr': lr_small})
        if self.coeff_type != 'none' and t.coeff:
            groups.append({'params': self.coeffs.parameters(), 'lr': lr_large})
        if 'fix' not in self.cfg.model.basis_type and self.cfg.model.basis_type != 'none' and t.basis:
            groups.append({'params': self.basises.parameters(), 'lr': lr_large})
        if 'reconstruction' in self.cfg.defaults.mode and t.renderModule:
            groups.append({'params': self.renderModule.parameters(), 'lr': lr_small})
        return groups

    def set_optimizable(self, items, statue):
        """FactorFields.py:556-567, quirks preserved: 'coeff' toggles the bases, 'proj'/'renderer' set a module attribute."""
        for item in items:
            if item == 'basis' and self.cfg.model.basis_type != 'none':
                for p in self.basises:
                    p.requires_grad = statue
            elif item == 'coeff' and self.cfg.model.coeff_type != 'none':
                for p in self.basises:
                    p.requires_grad = statue
            elif item == 'proj':
                self.linear_mat.requires_grad = statue
            elif item == 'renderer':
                self.renderModule.requires_grad = statue

    def TV_loss(self, reg):
        total = 0
        for idx in range(len(self.basises)):
            total = total + reg(self.basises[idx]) * 1e-2
        return total

    # ---- sampling -----------------------------------------------------------------------------------
    def _jitter(self, n_rays, is_train):
        """The reference draws ONE uniform per ray with torch.rand_like on the CPU generator (FactorFields.py:593-595)
        and ships the full [rays, samples] tensor to the device; we draw the same numbers and ship [rays]."""
        if not is_train:
            return None
        return torch.rand(n_rays, 1)[:, 0].to(self.device)

    def _sampler_desc(self, N_samples, with_alpha, alpha_thres=0.5, mode=ops.SAMPLE_BOUNDED, z_table=None, alpha_outside=False):
        return ops.make_sampler_desc(self.aabb[:, :self.in_dim], self.stepSize, N_samples,
                                     alpha=self.alphaMask if with_alpha else None, alpha_thres=alpha_thres,
                                     mode=mode, z_table=z_table, bg_len=getattr(self, 'bg_len', 0.0), alpha_outside=alpha_outside)

    def sample_point(self, rays_o, rays_d, is_train=True, N_samples=-1):
        """FactorFields.py:586-602 -> (rays_pts [R,S,3], interpx [R,S], ~mask_outbbox [R,S])"""
        N_samples = N_samples if N_samples > 0 else self.nSamples
        rays = torch.cat([rays_o, rays_d], -1).to(self.device).float()
        jitter = self._jitter(rays.shape[0], is_train)
        mask, z = ops.sample_dense(self._sampler_desc(N_samples, False), rays, jitter)
        pts = rays[:, None, :3] + rays[:, None, 3:6] * z[..., None]
        return pts, z, mask

    def _z_uniform(self, n, is_train):
        """The per-SAMPLE uniforms of sample_point_ndc (:579, torch.rand_like on a [1,S] row) and sample_point_unbound
        (:612); drawn from the CPU generator so that runs are reproducible against the reference's CPU path."""
        return torch.rand(1, n)[0] if is_train else None

    def _z_table_host(self, kind, N_samples, is_train):
        """The interpx row shared by all rays, evaluated on the host like the reference's CPU tensors.
        kind 'ndc' (FactorFields.py:577-580): linspace(near, far, N) + uniform * (far - near) / N when training.
        kind 'unbound' (:607-623): 3/4 of the samples linear in [0,2], 1/4 in inverse depth out to 32; a uniform point
        (training) or the midpoint (evaluation) of every bin."""
        if kind == 'ndc':
            near, far = self.cfg.dataset.near_far
            interpx = torch.linspace(near, far, N_samples).unsqueeze(0)
            u = self._z_uniform(N_samples, is_train)
            if u is not None:
                interpx += u[None] * ((far - near) / N_samples)
            return interpx[0].contiguous()
        N_inner, N_outer = 3 * N_samples // 4, N_samples // 4
        b_inner = torch.linspace(0, 2, N_inner + 1)
        b_outer = 2 / torch.linspace(1, 1 / 16, N_outer + 1)
        rng = self._z_uniform(N_inner + N_outer, is_train)
        if rng is not None:
            interpx = torch.cat([b_inner[1:] * rng[:N_inner] + b_inner[:-1] * (1 - rng[:N_inner]),
                                 b_outer[1:] * rng[N_inner:] + b_outer[:-1] * (1 - rng[N_inner:])])
        else:
            interpx = torch.cat([(b_inner[1:] + b_inner[:-1]) * 0.5, (b_outer[1:] + b_outer[:-1]) * 0.5])
        return interpx.contiguous()

    def _z_table(self, kind, N_samples, is_train):
        """Device copy of the interpx row; a caller that replays a captured graph (train.TrainStep) parks a static buffer in
        `_z_static` and refreshes it itself before every replay."""
        static = getattr(self, '_z_static', None)
        if static is not None:
            return static
        return self._z_table_host(kind, N_samples, is_train).to(self.device)

    def _z_table_ndc(self, N_samples, is_train):
        return self._z_table('ndc', N_samples, is_train)

    def _z_table_unbound(self, N_samples, is_train):
        return self._z_table('unbound', N_samples, is_train)

    def sample_point_ndc(self, rays_o, rays_d, is_train=True, N_samples=-1):
        """FactorFields.py:575-584 -> (rays_pts [R,S,3], interpx [1,S], ~mask_outbbox [R,S])"""
        N_samples = N_samples if N_samples > 0 else self.nSamples
        rays = torch.cat([rays_o, rays_d], -1).to(self.device).float()
        zt = self._z_table_ndc(N_samples, is_train)
        mask, _, pts = ops.sample_dense(self._sampler_desc(N_samples, False, mode=ops.SAMPLE_NDC, z_table=zt), rays, None,
                                        want_z=False, want_pts=True)
        return pts, zt[None], mask

    def sample_point_unbound(self, rays_o, rays_d, is_train=True, N_samples=-1):
        """FactorFields.py:604-633 -> (contracted rays_pts [R,S,3], interpx [1,S], inner_mask [R,S])"""
        N_samples = N_samples if N_samples > 0 else self.nSamples
        rays = torch.cat([rays_o, rays_d], -1).to(self.device).float()
        zt = self._z_table_unbound(N_samples, is_train)
        mask, _, pts = ops.sample_dense(self._sampler_desc(zt.numel(), False, mode=ops.SAMPLE_UNBOUND, z_table=zt), rays, None,
                                        want_z=False, want_pts=True)
        return pts, zt[None], mask

    def normalize_coord(self, xyz_sampled):
        """FactorFields.py:635-637"""
        invaabbSize = 2.0 / (self.aabb[1] - self.aabb[0])
        return (xyz_sampled - self.aabb[0]) * invaabbSize - 1

    def _cdesc(self, white_bg=True, white_bg_dev=None):
        r = self.cfg.renderer
        return ops.make_composite_desc(r.density_shift, r.fea2denseAct, r.distance_scale, r.rayMarch_weight_thres, white_bg, white_bg_dev)

    def basis2density(self, density_features):
        """FactorFields.py:639-643"""
        # API-only helper (differentiable, for callers' own code); forward() applies the activation inside the
        # composite kernels (composite.cu: density_act).
        shift = self.cfg.renderer.density_shift
        if self.cfg.renderer.fea2denseAct == 'softplus':
            return torch.nn.functional.softplus(density_features + shift)
        return torch.relu(density_features + shift)

    # ---- checkpointing (FactorFields.py:645-691) ------------------------------------------------------
    @torch.no_grad()
    def cal_mean_coef(self, state_dict):
        if 'grid' in self.coeff_type or 'mlp' in self.coeff_type:
            keys = [k for k in state_dict.keys() if 'coeffs.0' in k]
            for key in keys:
                average = torch.zeros_like(state_dict[key])
                for i in range(self.n_scene):
                    item = key.replace('0', f'{i}', 1)
                    average += state_dict[item]
                    state_dict.pop(item, None)
                average /= self.n_scene
                state_dict[key] = average
        elif 'vec' in self.coeff_type:
            state_dict['coeffs.0'] = torch.mean(state_dict['coeffs.0'], dim=-1, keepdim=True)
        elif 'cp' in self.coeff_type or 'vm' in self.coeff_type:
            for i in range(3):
                state_dict[f'coeffs.{i}'] = torch.mean(state_dict[f'coeffs.{i}'], dim=-1, keepdim=True)
        return state_dict

    def save(self, path):
        sd = {k: v.contiguous() for k, v in self.state_dict().items()}   # reference (channel-first) storage order on disk
        ckpt = {'state_dict': sd, 'cfg': self.cfg}
        if self.alphaMask is not None:
            alpha_volume = self.alphaMask.alpha_volume.bool().cpu().numpy()
            ckpt.update({'alphaMask.shape': alpha_volume.shape})
            ckpt.update({'alphaMask.mask': np.packbits(alpha_volume.reshape(-1))})
            ckpt.update({'alphaMask.aabb': self.alphaMask.aabb.cpu()})
        if 'reconstruction' in self.cfg.defaults.mode:
            ckpt['state_dict'] = self.cal_mean_coef(ckpt['state_dict'])
        torch.save(ckpt, path)

    def load(self, ckpt):
        if 'alphaMask.aabb' in ckpt.keys():
            length = np.prod(ckpt['alphaMask.shape'])
            alpha_volume = torch.from_numpy(np.unpackbits(ckpt['alphaMask.mask'])[:length].reshape(ckpt['alphaMask.shape']))
            self.alphaMask = AlphaGridMask(self.device, ckpt['alphaMask.aabb'].to(self.device), alpha_volume.float().to(self.device))
        self.load_state_dict(ckpt['state_dict'])
        volumeSize = N_to_reso(self.cfg.training.volume_resoFinal ** self.in_dim, self.aabb)
        self.update_renderParams(volumeSize)

    # ---- render parameters (FactorFields.py:693-708) ------------------------------------------------
    def update_renderParams(self, gridSize):
        # The reference evaluates these few scalars with torch on self.device; a 3-element torch.mean rounds
        # differently on CUDA and on CPU, and stepSize feeds bit-exact sampling decisions, so they are evaluated
        # on the host (= the reference's CPU path, which is what the oracle is pinned to).
        self.aabbSize = self.aabb[1] - self.aabb[0]
        self.gridSize = torch.LongTensor(gridSize).to(self.device)
        size_h = self.aabbSize.cpu()
        units = size_h / (torch.LongTensor(gridSize) - 1)
        step_h = torch.mean(units) * self.cfg.renderer.step_ratio
        self.stepSize = step_h.to(self.device)
        aabbDiag = torch.sqrt(torch.sum(torch.square(size_h)))
        self.nSamples = int((aabbDiag / step_h).item()) + 1

    @torch.no_grad()
    def upsample_volume_grid(self, res_target):
        self.update_renderParams(res_target)
        if self.cfg.dataset.dataset_name == 'google_objs' and self.n_scene == 1 and self.cfg.model.coeff_type == 'grid':
            up = torch.nn.functional.interpolate(self.coeffs[0].data, size=None, scale_factor=1.3, align_corners=True, mode='trilinear')
            self.coeffs = torch.nn.ParameterList([_factor_param(up)])
            self._plans = {}

    # ---- alpha mask maintenance (FactorFields.py:710-841) ---------------------------------------------
    def compute_alpha(self, xyz_locs, length=1):
        shape = xyz_locs.shape[:-1]
        xyz = xyz_locs.reshape(-1, 3)
        if self.alphaMask is not None:
            alpha_mask = self.alphaMask.sample_alpha(xyz) > 0
        else:
            alpha_mask = torch.ones_like(xyz[:, 0], dtype=bool)
        alpha = torch.zeros(xyz.shape[0], device=xyz.device)
        if alpha_mask.any():
            feats, _ = self.get_coding(xyz[alpha_mask])
            feat = self.linear_mat(feats, is_train=False)
            alpha[alpha_mask] = ops.density_alpha(self._cdesc(), feat, float(length))
        return alpha.view(shape)

    @torch.no_grad()
    def getDenseAlpha(self, gridSize=None, times=16, sharded=True, chunk_points=1 << 22):
        """FactorFields.py:730-755: alpha of the voxel-centre lattice, averaged over `times` jittered evaluations.

        GPU-side formulation: the lattice [Z, Y, X, 3] is built on the device by broadcasting, and every round is evaluated
        in a few launches over slabs of up to `chunk_points` points (157 M field queries per event at 214^3 used to be 16 x 214
        Python iterations with a CPU draw, a host->device copy and a launch chain each).  The jitter keeps the reference's
        random stream: one `torch.rand(Z, Y, X, 3)` per round on the CPU generator yields exactly the numbers of the
        reference's Z per-slice draws (same generator, same order), so the volume is bit-identical to the per-slice loop.
        Under torch.distributed (sharded=True) rank r evaluates a contiguous block of slices and the partial volumes are summed
        with one all-reduce; every rank still draws every round's full jitter, keeping the ranks' generators in step."""
        rank, world = 0, 1
        if sharded and torch.distributed.is_available() and torch.distributed.is_initialized():
            rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
        gridSize = [int(g) for g in (self.gridSize.tolist() if gridSize is None else gridSize)]
        X, Y, Z = gridSize
        dev = self.device
        lo, hi = self.inward_aabb[0], self.inward_aabb[1]
        cells = torch.LongTensor(gridSize) - 1
        units = (hi - lo) / cells.to(dev)
        half = 0.5 / cells.float()
        length = torch.mean(units.cpu()) * self.cfg.renderer.distance_scale     # host evaluation, see update_renderParams
        # voxel centres per axis (host linspace, as the reference's), blended with the box corners on the device
        s = [torch.linspace(float(half[k]), float(1 - half[k]), gridSize[k]).to(dev) for k in range(3)]
        ax = [lo[k] * (1 - s[k]) + hi[k] * s[k] for k in range(3)]
        dense_xyz = torch.stack([ax[0].view(1, 1, X).expand(Z, Y, X), ax[1].view(1, Y, 1).expand(Z, Y, X),
                                 ax[2].view(Z, 1, 1).expand(Z, Y, X)], -1).contiguous()
        alpha = torch.zeros((Z, Y, X), device=dev)
        base, rem = divmod(Z, world)
        z_lo = rank * base + min(rank, rem)
        z_hi = z_lo + base + (1 if rank < rem else 0)
        slab = max(1, int(chunk_points) // max(1, X * Y))
        amp = units / 2 * 1.2
        for _ in range(times):
            shift = torch.rand(Z, Y, X, 3) if times > 1 else None          # drawn on every rank: the same stream everywhere
            for z0 in range(z_lo, z_hi, slab):
                z1 = min(z0 + slab, z_hi)
                pts = dense_xyz[z0:z1]
                if shift is not None:
                    pts = pts + (shift[z0:z1].to(dev, non_blocking=True) * 2 - 1) * amp
                alpha[z0:z1] += self.compute_alpha(pts.reshape(-1, 3), length).view(z1 - z0, Y, X)
        if world > 1:
            torch.distributed.all_reduce(alpha, op=torch.distributed.ReduceOp.SUM)
        return alpha / times, dense_xyz

    @torch.no_grad()
    def updateAlphaMask(self, gridSize=(200, 200, 200), is_update_alphaMask=False):
        alpha, dense_xyz = self.getDenseAlpha(gridSize)
        ks = 3
        alpha = alpha.clamp(0, 1)[None, None]
        alpha = torch.nn.functional.max_pool3d(alpha, kernel_size=ks, padding=ks // 2, stride=1).view(gridSize[::-1])
        min_size = np.mean(alpha.shape[-3:]).item()
        alphaMask_thres = self.cfg.renderer.alphaMask_thres if is_update_alphaMask else 0.08
        if self.is_unbound:
            alphaMask_thres = 0.04
            alpha = (alpha >= alphaMask_thres).float()
        else:
            keep = remove_small_objects(alpha.cpu().numpy() >= alphaMask_thres, min_size=min_size, connectivity=1)
            alpha = torch.FloatTensor(keep).to(self.device)
        if is_update_alphaMask:
            self.alphaMask = AlphaGridMask(self.device, self.inward_aabb, alpha)
        valid_xyz = dense_xyz[alpha > 0.5]
        xyz_min, xyz_max = valid_xyz.amin(0), valid_xyz.amax(0)
        if not self.is_unbound:
            pad = (xyz_max - xyz_min) / 20
            xyz_min -= pad
            xyz_max += pad
        return torch.stack((xyz_min, xyz_max))

    @torch.no_grad()
    def shrink(self, new_aabb):
        """FactorFields.py:795-809 — re-initialises the factors at the new box (the reference's behaviour)."""
        self.setup_params(new_aabb.tolist())
        if self.cfg.model.coeff_type != 'none':
            del self.coeffs
            self.coeffs = self.init_coef()
        if self.cfg.model.basis_type != 'none':
            del self.basises
            self.basises = self.init_basis()
        self.aabb = self.inward_aabb = new_aabb
        self.cfg.dataset.aabb = self.aabb.tolist()
        self.update_renderParams(self.gridSize.tolist())
        self._plans = {}

    @torch.no_grad()
    def filtering_rays(self, all_rays, all_rgbs, N_samples=256, chunk=10240 * 5, bbox_only=False):
        """FactorFields.py:811-841 — keeps the rays whose samples touch the alpha mask (kernel: per-ray valid counts)."""
        N = int(torch.tensor(all_rays.shape[:-1]).prod())
        kept = 0
        for idx_chunk in torch.split(torch.arange(N), chunk):
            rays_chunk = all_rays[idx_chunk].to(self.device).float().contiguous()
            if bbox_only:
                rays_o, rays_d = rays_chunk[..., :3], rays_chunk[..., 3:6]
                vec = torch.where(rays_d == 0, torch.full_like(rays_d, 1e-6), rays_d)
                rate_a = (self.aabb[1] - rays_o) / vec
                rate_b = (self.aabb[0] - rays_o) / vec
                mask_inbbox = torch.maximum(rate_a, rate_b).amin(-1) > torch.minimum(rate_a, rate_b).amax(-1)
            else:
                desc = self._sampler_desc(N_samples, True, alpha_thres=0.0, alpha_outside=True)   # :832-833 tests every sample
                counts = torch.empty(rays_chunk.shape[0], device=self.device, dtype=torch.int32)
                nv.check(nv.lib().ffb_sample_count(C.byref(desc), nv.ptr(rays_chunk), None, C.c_int64(rays_chunk.shape[0]),
                                                   nv.i32p(counts), None, nv.stream()))
                mask_inbbox = counts > 0
            length = int(mask_inbbox.sum())
            sel = mask_inbbox.cpu()
            all_rays[kept:kept + length], all_rgbs[kept:kept + length] = rays_chunk[mask_inbbox].cpu(), all_rgbs[idx_chunk][sel]
            kept += length
        return all_rays[:kept], all_rgbs[:kept]

    # ---- the render step (FactorFields.py:843-898) ----------------------------------------------------
    def forward(self, rays_chunk, white_bg=True, is_train=False, ndc_ray=False, N_samples=-1):
        if not rays_chunk.is_cuda:
            raise RuntimeError('rays must be on the CUDA device (renderer.render_ray does the host->device copy)')
        N_samples = N_samples if N_samples > 0 else self.nSamples
        rays = rays_chunk[:, :6].contiguous().float()
        lazy = bool(getattr(self, 'lazy_counts', False)) and not (self._coeff_is_mlp() or self._basis_is_mlp())
        with_alpha = self.alphaMask is not None
        if self.is_unbound:     # :847-850
            zt = self._z_table_unbound(N_samples, is_train)
            N_samples = zt.numel()   # 3N//4 + N//4 can be below N
            samp = ops.sample_compact(self._sampler_desc(N_samples, with_alpha, mode=ops.SAMPLE_UNBOUND, z_table=zt), rays, None, lazy=lazy)
        elif ndc_ray:           # :851-857; the appearance MLP sees unit view directions
            zt = self._z_table_ndc(N_samples, is_train)
            samp = ops.sample_compact(self._sampler_desc(N_samples, with_alpha, mode=ops.SAMPLE_NDC, z_table=zt), rays, None, lazy=lazy)
            samp['rays'] = torch.cat([rays[:, :3], rays[:, 3:6] / torch.norm(rays[:, 3:6], dim=-1, keepdim=True)], -1).contiguous()
        else:                   # :858-861
            jitter = self._jitter(rays.shape[0], is_train)
            samp = ops.sample_compact(self._sampler_desc(N_samples, with_alpha), rays, jitter, lazy=lazy)
        self.last_stats = {'n_valid': samp['n_valid'], 'n_candidates': rays.shape[0] * N_samples}

        # :890 `white_bg or (is_train and torch.rand((1,)) < 0.5)`.  A caller that replays a captured graph (train.TrainStep)
        # parks a device flag in `_white_bg_static` and refreshes it per step, so the coin flip is not frozen at capture.
        bg_dev = getattr(self, '_white_bg_static', None)
        if bg_dev is None:
            white_bg = bool(white_bg or (is_train and torch.rand((1,)) < 0.5))
        width = sum(self.cfg.model.basis_dims)
        if lazy or samp['n_valid'] > 0:
            feats, coeffs = self.get_coding(samp['xyz'], samp['n_dev'])
            feat = self.linear_mat(feats, is_train=is_train, n_dev=samp['n_dev'])
        else:
            coeffs = torch.zeros((1, width), device=rays.device)
            feat = torch.zeros((0, self.cfg.model.out_dim), device=rays.device)
        params, has_bias = self.renderModule._flat()
        rgb_map, depth_map, acc, weight, app_idx, n_app = ops.RenderComposite.apply(
            feat, samp, self._cdesc(white_bg, bg_dev), self.renderModule.viewpe, self.renderModule.feape, has_bias, *params)
        self.last_stats['n_app'] = n_app if lazy else int(n_app)
        self.last_aux = dict(samp=samp, weight=weight, app_idx=app_idx, acc=acc)
        return rgb_map, depth_map, coeffs
