"""Helpers for the -m gpu parity tests: build the product model from a golden case."""
import json

import numpy as np
import torch

import ffb200
from ffb200.models.FactorFields import FactorFields
from tests import helpers as H


def make_cfg(g):
    cfgname, ov, aabb = H.load_ref_cfg(g)
    cfg = ffb200.load_cfg(cfgname, [f'{k}={json.dumps(v)}' for k, v in ov.items()])
    if cfgname == 'image_set.yaml':
        aabb = [[int(v) for v in r] for r in aabb]
    cfg.dataset.aabb = aabb
    return cfg


def build_model(g, device='cuda'):
    cfg = make_cfg(g)
    m = FactorFields(cfg, device)
    sd = {k[len('param.'):]: torch.from_numpy(np.ascontiguousarray(v)) for k, v in g.items() if k.startswith('param.')}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    if 'scene_idx' in g:
        m.scene_idx = int(g['scene_idx'])
        m._plans = {}
    return cfg, m


def t(a, device='cuda'):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def npy(x):
    return x.detach().cpu().contiguous().numpy()
