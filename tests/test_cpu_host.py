"""CPU-only checks: the C-ABI library loads and exports every declared symbol, the host logic (config, shape
derivation) matches the reference's facts, and the product refuses to run without CUDA."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import ffb200
from tests import helpers as H

ROOT = H.ROOT


def test_library_exports_every_declared_symbol():
    from ffb200 import native as nv
    hdr = open(os.path.join(ROOT, 'include', 'ffb200.h')).read()
    names = set(re.findall(r'^\s*(?:int|uint64_t|const char\*)\s+(ffb_\w+)\s*\(', hdr, re.M))
    assert len(names) >= 30
    lib = nv.lib()
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.ffb_abi_version() == 1


def test_struct_layouts_match_header():
    """sizeof() of the ctypes mirrors == sizeof() of the C structs (compiled with gcc from the header)."""
    from ffb200 import native as nv
    src = '#include <stdio.h>\n#include "ffb200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(ffb_gather_op), sizeof(ffb_term),' \
          ' sizeof(ffb_field_desc), sizeof(ffb_sampler_desc), sizeof(ffb_composite_desc));return 0;}'
    exe = '/tmp/ffb_sizeof'
    subprocess.run(['gcc', '-x', 'c', '-', '-I', os.path.join(ROOT, 'include'), '-o', exe], input=src.encode(), check=True)
    sizes = [int(v) for v in subprocess.run([exe], capture_output=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(nv.GatherOp), C.sizeof(nv.Term), C.sizeof(nv.FieldDesc), C.sizeof(nv.SamplerDesc), C.sizeof(nv.CompositeDesc)]


def test_no_cpu_fallback():
    from ffb200.models.FactorFields import FactorFields
    cfg = ffb200.load_cfg('nerf.yaml')
    cfg.dataset.aabb = [[-1., -1., -1.], [1., 1., 1.]]
    with pytest.raises(RuntimeError):
        FactorFields(cfg, 'cpu')
    from ffb200 import ops
    with pytest.raises(RuntimeError):
        ops.grid_mapping(torch.zeros(3, 3), torch.ones(2), torch.tensor([[0., 0, 0], [1, 1, 1]]))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'factor-fields_b200')
    pat = re.compile(r'^\s*(from|import)\s+oracle|ff_oracle|torch_port|import_module\([\'"]oracle', re.M)
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                assert not pat.search(open(os.path.join(dp, f)).read()), f


def test_config_loader_and_overrides():
    cfg = ffb200.load_cfg('nerf.yaml', ['model.basis_type=vm', 'training.n_iters=100', 'model.basis_dims=[18]'])
    assert cfg.renderer.rayMarch_weight_thres == 1e-3 and isinstance(cfg.renderer.rayMarch_weight_thres, float)
    assert cfg.model.basis_type == 'vm' and cfg.training.n_iters == 100 and cfg.model.basis_dims == [18]
    assert cfg.training.lr_large == 0.02 and cfg.model.T_basis == 0      # inherited from defaults.yaml
    # dot-list values carry YAML semantics like OmegaConf.from_cli (train_per_scene.py:245): booleans, null, exponents
    cfg = ffb200.load_cfg('image_set.yaml', ['model.with_dropout=false', 'defaults.ckpt=null', 'renderer.alphaMask_thres=4e-2',
                                             'dataset.is_unbound=true'])
    assert cfg.model.with_dropout is False and cfg.defaults.ckpt is None and cfg.renderer.alphaMask_thres == 0.04
    assert cfg.dataset.is_unbound is True
    assert ffb200.load_cfg('360_v2.yaml').dataset.is_unbound is True and ffb200.load_cfg('360_v2.yaml').renderer.fea2denseAct == 'relu'


@pytest.mark.parametrize('name', H.field_cases())
def test_shape_logic_matches_reference_facts(name):
    from ffb200.models.FactorFields import field_shapes
    g = H.golden('field_' + name)
    cfgname, ov, aabb = H.load_ref_cfg(g)
    cfg = ffb200.load_cfg(cfgname, [f'{k}={json.dumps(v)}' for k, v in ov.items()])
    if cfgname == 'image_set.yaml':
        aabb = [[int(v) for v in r] for r in aabb]
    sh = field_shapes(cfg, aabb)
    assert sh['in_dim'] == int(g['fact.in_dim'])
    assert np.array_equal(sh['freq_bands'].numpy(), g['fact.freq_bands'])
    assert np.array_equal(sh['aabb'].numpy(), g['fact.aabb'])
    if not np.isnan(np.array(sh['basis_reso'], float)).any():
        assert list(sh['basis_reso']) == list(g['fact.basis_reso'])
    if 'fact.coeff_reso' in g:
        assert [int(v) for v in sh['coeff_reso']] == list(g['fact.coeff_reso'])


def test_resolution_helpers():
    from ffb200.utils import N_to_reso, N_to_vm_reso, cal_n_samples, SimpleSampler
    box = torch.tensor([[-1.2, -0.7, -1.0], [1.3, 0.9, 0.8]])
    # expected values: the reference's utils.N_to_reso / N_to_vm_reso on the same box (recorded in the dev container)
    assert N_to_reso(128 ** 3, box) == [166, 106, 119] and N_to_reso(300 ** 3, box) == [388, 249, 280]
    assert cal_n_samples([128, 128, 128], 0.5) == 443
    np.random.seed(3)
    s = SimpleSampler(100, 32)
    ids = [s.nextids() for _ in range(4)]
    assert all(len(i) == 32 for i in ids[:3])
    assert N_to_vm_reso(64 ** 3, box) == [379, 243, 273]


def test_math_header_host_build_matches_oracle():
    """ffb_math.h is host+device code: compile it with g++ and check the sampler decisions bit-exactly vs the oracle."""
    from oracle import ff_oracle as O
    shim = r'''
#include "ffb_math.h"
extern "C" void run(const float* rays, const float* jit, int R, int S, const float* lo, const float* hi, float step, unsigned char* mask, float* z) {
  for (int r = 0; r < R; ++r) {
    float tmin = ffb::ray_tmin(rays + r * 6, rays + r * 6 + 3, lo, hi);
    for (int s = 0; s < S; ++s) {
      float p[3];
      float t = ffb::sample_t(tmin, step, s, jit ? jit[r] : 0.f, jit != nullptr);
      mask[r * S + s] = ffb::sample_pos(rays + r * 6, rays + r * 6 + 3, t, lo, hi, p);
      z[r * S + s] = t;
    }
  }
}
extern "C" float mapc(float x, float lo, float scale, int mode) { return ffb::map_coord(x, lo, scale, mode, nullptr); }
extern "C" void run_unbound(const float* rays, const float* zt, int R, int S, float bg, unsigned char* inner, float* pts) {
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s)
      inner[r * S + s] = ffb::sample_pos_unbound(rays + r * 6, rays + r * 6 + 3, zt[s], bg, pts + (size_t)(r * S + s) * 3);
}
'''
    so = '/tmp/ffb_math_shim.so'
    subprocess.run(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-x', 'c++', '-', '-I', os.path.join(ROOT, 'factor-fields_b200', 'csrc'),
                    '-o', so], input=shim.encode(), check=True)
    lib = C.CDLL(so)
    g = H.golden('sampler_nerf')
    rays = np.ascontiguousarray(g['rays'][:256], np.float32)
    jit = np.ascontiguousarray(g['jitter'][:256], np.float32)
    R, S = 256, 443
    mask = np.zeros((R, S), np.uint8)
    z = np.zeros((R, S), np.float32)
    lo, hi = np.ascontiguousarray(g['aabb'][0]), np.ascontiguousarray(g['aabb'][1])
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.run(P(rays), P(jit), R, S, P(lo), P(hi), C.c_float(float(g['stepSize'])), P(mask), P(z))
    _, z_ref, inner = O.sample_point(g['aabb'], g['stepSize'], rays[:, :3], rays[:, 3:], S, jit)
    assert np.array_equal(mask.astype(bool), inner) and np.array_equal(z, z_ref)
    lib.mapc.restype = C.c_float
    xs = np.random.RandomState(0).uniform(-1.2, 1.3, 2000).astype(np.float32)
    for mode_id, mode in enumerate(['sawtooth', 'triangle']):
        ref = O.grid_mapping(xs[:, None], np.array([3.1], np.float32), np.array([[-1.2], [1.3]], np.float32), mode)[:, 0, 0]
        scale = np.float32(np.float32(2.5) / np.float32(3.1))
        got = np.array([lib.mapc(C.c_float(float(v)), C.c_float(-1.2), C.c_float(float(scale)), mode_id) for v in xs], np.float32)
        assert np.array_equal(got, ref), mode
    # unbounded scenes: inf-norm contraction (FactorFields.py:625-633) bit-exact vs the oracle, on the golden case's rays / interpx
    gu = H.golden('render_unbound_train')
    rays_u = np.ascontiguousarray(gu['rays'], np.float32)
    zt = np.ascontiguousarray(gu['z'][0], np.float32)
    Ru, Su = rays_u.shape[0], zt.shape[0]
    inner_u = np.zeros((Ru, Su), np.uint8)
    pts_u = np.zeros((Ru, Su, 3), np.float32)
    lib.run_unbound(P(rays_u), P(zt), Ru, Su, C.c_float(float(gu['bg_len'])), P(inner_u), P(pts_u))
    pts_ref, _, inner_ref = O.sample_point_unbound(float(gu['bg_len']), zt, rays_u[:, :3], rays_u[:, 3:])
    assert np.array_equal(inner_u.astype(bool), inner_ref) and np.array_equal(pts_u, pts_ref)
    assert np.array_equal(np.packbits(inner_ref), gu['inner_mask'])


def test_dct_dict_matches_reference():
    """Host-side DCT dictionary (FactorFields.py:36-71) that initialises every grid basis: equal to the reference's values."""
    from ffb200.models.FactorFields import dct_dict
    g = H.golden('api')
    assert H.rel_err(dct_dict(3, 12, n_selete=5, dim=2).numpy(), g['dct2']) < 1e-6
    assert H.rel_err(dct_dict(2, 7, n_selete=4, dim=3).numpy(), g['dct3']) < 1e-6
