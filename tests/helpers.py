"""Shared helpers for the parity tests (test infrastructure)."""
import glob
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))


def field_cases():
    return sorted(os.path.basename(p)[len('field_'):-4] for p in glob.glob(os.path.join(GOLDEN, 'field_*.npz')))


def load_ref_cfg(g):
    """(cfg yaml name, overrides dict, aabb list) of a golden case."""
    return str(g['cfgname']), json.loads(str(g['overrides'])), g['aabb_cfg'].tolist()


def ref_mode(cfgname):
    return {'nerf.yaml': 'reconstruction', 'sdf.yaml': 'sdf', 'image.yaml': 'image', 'image_set.yaml': 'images'}[cfgname]


def model_defaults(cfgname):
    """coeff/basis types and modes of the reference yaml files (configs/*.yaml) the golden cases start from."""
    base = dict(coeff_type='grid', basis_type='grid', basis_mapping='sawtooth', coef_mode='bilinear', basis_mode='bilinear')
    if cfgname == 'image.yaml':
        base.update(coef_mode='nearest', basis_mode='nearest')
    return base


def oracle_spec(g):
    cfgname, ov, _ = load_ref_cfg(g)
    md = model_defaults(cfgname)
    for k, v in ov.items():
        sec, key = k.split('.')
        if sec == 'model' and key in md:
            md[key] = v
    spec = dict(mode=ov.get('defaults.mode', ref_mode(cfgname)), in_dim=int(g['fact.in_dim']), aabb=g['fact.aabb'],
                freq_bands=g['fact.freq_bands'], basis_dims=g['fact.basis_dims'].tolist(), **md)
    if 'n_scene' in g:
        spec.update(n_scene=int(g['n_scene']), scene_idx=int(g['scene_idx']))
    return spec


def _layers(g, prefix, names=('backbone', 'mlp')):
    """[(W, b)] from 'param.<prefix>.<backbone|mlp>.<i>.weight/bias' entries."""
    out, i = [], 0
    while True:
        key = None
        for nm in names:
            k = f'{prefix}.{nm}.{i}.weight' if nm else f'{prefix}.{i}.weight'
            if k in g:
                key = k
        if key is None:
            break
        b = key[:-len('weight')] + 'bias'
        out.append((g[key], g[b] if b in g else None))
        i += 1
    return out


def oracle_params(g):
    params = {'coeffs': [], 'basises': []}
    for kind in ('coeffs', 'basises'):
        i = 0
        while True:
            if f'param.{kind}.{i}' in g:
                params[kind].append(g[f'param.{kind}.{i}'])
            elif f'param.{kind}.{i}.backbone.0.weight' in g:
                params[kind].append(_layers(g, f'param.{kind}.{i}'))
            else:
                break
            i += 1
    return params


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny): 'relative to the largest reference magnitude' (SURVEY §4)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)) if a.size else 0.0
