"""GPU parity of sampling/compaction (K4), MLPs (K3), composite (K5) and the full forward()/backward against the
golden vectors of the unmodified reference and against the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_sampler_bit_exact_through_c_abi():
    """Host-buffer C-ABI entry point, no torch involved: packed in-box masks must equal the reference's."""
    from ffb200 import native as nv
    g = H.golden('sampler_nerf')
    rays = np.ascontiguousarray(g['rays'], np.float32)
    R = rays.shape[0]
    for S, jitter, key in [(443, np.ascontiguousarray(g['jitter'], np.float32), 'inner_train'), (int(g['nSamples']), None, 'inner_eval')]:
        d = nv.SamplerDesc()
        for k in range(3):
            d.aabb_min[k], d.aabb_max[k] = float(g['aabb'][0, k]), float(g['aabb'][1, k])
        d.step_size = float(g['stepSize'])
        d.n_samples = S
        d.alpha_volume = 0
        mask = np.zeros((R, S), np.uint8)
        z = np.zeros((R, S), np.float32)
        nv.check(nv.lib().ffb_sample_dense_host(C.byref(d), rays.ctypes.data_as(C.c_void_p),
                                                jitter.ctypes.data_as(C.c_void_p) if jitter is not None else None, C.c_int64(R),
                                                mask.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p)))
        assert np.array_equal(np.packbits(mask.astype(bool)), g[key])
        if jitter is not None:
            assert np.array_equal(z[:, 0], g['z_train_first']) and np.array_equal(z[:, -1], g['z_train_last'])
            assert np.array_equal(mask.sum(-1), g['counts_train'])


def test_mlps_golden():
    from ffb200.models.FactorFields import MLPMixer, MLPRender_Fea
    from tests import gpu_helpers as G
    g = H.golden('mlp')
    for tag in ('lm_nerf', 'lm_sdf', 'mlpC', 'deep'):
        i, o, L, hdim, pe = [int(v) for v in g[f'{tag}.cfg']]
        mm = MLPMixer(i, o, num_layers=L, hidden_dim=hdim, pe=pe).cuda()
        mm.load_state_dict({k[len(tag) + 7:]: torch.from_numpy(v) for k, v in g.items() if k.startswith(f'{tag}.param.')})
        x = G.t(g[f'{tag}.x']).requires_grad_(True)
        y = mm(x)
        assert H.rel_err(G.npy(y), g[f'{tag}.y']) < 2e-5, tag
        grads = torch.autograd.grad((y * G.t(g[f'{tag}.G'])).sum(), [x] + list(mm.parameters()))
        assert H.rel_err(G.npy(grads[0]), g[f'{tag}.gx']) < 5e-5, tag
        for (n, p), gr in zip(mm.named_parameters(), grads[1:]):
            assert H.rel_err(G.npy(gr), g[f'{tag}.grad.{n}']) < 5e-5, (tag, n)
    rm = MLPRender_Fea(inChanel=31, num_layers=3, hidden_dim=128, viewpe=6, feape=2).cuda()
    rm.load_state_dict({k[len('rm.param.'):]: torch.from_numpy(v) for k, v in g.items() if k.startswith('rm.param.')})
    feat = G.t(g['rm.feat']).requires_grad_(True)
    y = rm(G.t(g['rm.vd']), feat)
    assert H.rel_err(G.npy(y), g['rm.y']) < 2e-5
    grads = torch.autograd.grad((y * G.t(g['rm.G'])).sum(), [feat] + list(rm.parameters()))
    assert H.rel_err(G.npy(grads[0]), g['rm.gfeat']) < 5e-5
    for (n, p), gr in zip(rm.named_parameters(), grads[1:]):
        assert H.rel_err(G.npy(gr), g[f'rm.grad.{n}']) < 5e-5, n


def _unpack(bits, shape):
    return np.unpackbits(bits)[:int(np.prod(shape))].reshape(shape).astype(bool)


@pytest.mark.parametrize('name', ['train', 'train_alpha', 'eval_alpha', 'ndc_train', 'ndc_eval_alpha', 'unbound_train',
                                  'unbound_eval_alpha'])
def test_render_golden(name):
    """forward(): bit-exact sample indices / counts; rgb, depth, coeffs, loss and every parameter gradient within 1e-4.
    Bounded (sample_point), NDC (sample_point_ndc) and unbounded (sample_point_unbound) scenes."""
    from ffb200.models.FactorFields import AlphaGridMask
    from ffb200.renderer import render_ray
    from tests import gpu_helpers as G
    g = H.golden('render_' + name)
    cfg, m = G.build_model(g)
    assert m.nSamples == int(g['fact.nSamples']) and float(m.stepSize) == float(g['fact.stepSize'])
    if 'alpha_volume' in g:
        m.alphaMask = AlphaGridMask('cuda', G.t(g['alpha_aabb']), G.t(g['alpha_volume']))
    is_train = bool(g['is_train'])
    S = int(g['N_samples'])
    R = g['rays'].shape[0]
    mode = str(g['mode']) if 'mode' in g else 'bounded'
    if is_train:                                                           # inject the reference's random numbers
        m._jitter = lambda n, tr: G.t(g['jitter']) if tr else None
        m._z_uniform = lambda n, tr: torch.from_numpy(g['jitter']) if tr else None
    rays_host = torch.from_numpy(g['rays'])                                # host rays: render_ray does the H2D copy
    out = render_ray(rays_host, m, chunk=4096, N_samples=S, ndc_ray=(mode == 'ndc'), white_bg=True, is_train=is_train, device='cuda')
    rgb_map, depth_map = out[0], out[1]
    aux = m.last_aux
    S = g['z'].shape[1]                                                    # unbounded: 3S//4 + S//4 samples
    # --- the public sample_point* methods: same masks / interpx / points as the reference's
    sampler = {'bounded': m.sample_point, 'ndc': m.sample_point_ndc, 'unbound': m.sample_point_unbound}[mode]
    pts, zz, inner = sampler(G.t(g['rays'][:, :3]), G.t(g['rays'][:, 3:6]), is_train=is_train, N_samples=int(g['N_samples']))
    assert np.array_equal(np.packbits(G.npy(inner)), g['inner_mask'])
    assert np.array_equal(np.broadcast_to(G.npy(zz), (R, S)), g['z'])
    if 'pts_sum' in g:
        assert np.allclose(G.npy(pts).astype(np.float64).sum((0, 1)), g['pts_sum'], rtol=1e-9, atol=1e-9)
    # --- bit-exact decisions
    valid_ref = _unpack(g['ray_valid'], (R, S))
    rr, ss = np.nonzero(valid_ref)
    assert aux['samp']['n_valid'] == int(g['n_valid'])
    assert np.array_equal(G.npy(aux['samp']['ray_id']), rr.astype(np.int32))
    assert np.array_equal(G.npy(aux['samp']['sample_id']), ss.astype(np.int32))
    assert np.array_equal(G.npy(aux['samp']['z']), g['z'][valid_ref])
    assert np.array_equal(G.npy(aux['samp']['counts']), valid_ref.sum(-1).astype(np.int32))
    # --- floating point
    w_ref = g['weight'][valid_ref]
    assert H.rel_err(G.npy(aux['weight']), w_ref) < 1e-4
    app_ref = _unpack(g['app_mask'], (R, S))[valid_ref]
    app = np.zeros(w_ref.shape[0], bool)
    app[G.npy(aux['app_idx'])] = True
    band = np.abs(w_ref - 1e-3) < 1e-6
    assert np.array_equal(app[~band], app_ref[~band])
    assert H.rel_err(G.npy(rgb_map), g['rgb_map']) < 1e-4
    assert H.rel_err(G.npy(depth_map), g['depth_map']) < 1e-4
    if is_train:
        assert H.rel_err(G.npy(out[2]), g['coeffs']) < 2e-5
    loss = torch.mean((rgb_map - G.t(g['target'])) ** 2)
    assert abs(float(loss) - float(g['loss'])) < 1e-5
    params = list(m.named_parameters())
    grads = torch.autograd.grad(loss, [p for _, p in params], allow_unused=True)
    for (n, p), gr in zip(params, grads):
        ref = g['grad.' + n]
        got = G.npy(gr) if gr is not None else np.zeros_like(ref)
        assert H.rel_err(got, ref) < 2e-4, (name, n)


def test_render_full_size_properties():
    """nerf.yaml shapes, 4096 rays x 443 samples: compaction order, counts, composite invariants, oracle spot check."""
    import ffb200
    from ffb200 import ops
    from ffb200.models.FactorFields import FactorFields
    from oracle import ff_oracle as O
    from tests.golden_rays import blender_like_rays
    cfg = ffb200.load_cfg('nerf.yaml')
    cfg.dataset.aabb = [[-1., -1., -1.], [1., 1., 1.]]
    torch.manual_seed(1)
    m = FactorFields(cfg, 'cuda')
    rays = blender_like_rays(4096, 3)
    jitter = np.random.RandomState(4).rand(4096).astype(np.float32)
    m._jitter = lambda n, tr: torch.from_numpy(jitter).cuda()
    rgb_map, depth, _ = m(torch.from_numpy(rays).cuda(), white_bg=True, is_train=True, N_samples=443)
    s = m.last_aux['samp']
    rid, sid = s['ray_id'].long(), s['sample_id'].long()
    key = rid * 443 + sid
    assert bool((key[1:] > key[:-1]).all())                      # strictly increasing row-major order
    off = s['offsets'].long()
    assert bool((off[1:] >= off[:-1]).all()) and int(off[-1]) == s['n_valid'] == key.numel()
    # oracle agrees on the mask, bit-exactly, at full size
    _, _, inner = O.sample_point(np.array(cfg.dataset.aabb, np.float32), np.float32(m.stepSize.item()), rays[:, :3], rays[:, 3:], 443, jitter)
    rr, ss = np.nonzero(inner)
    assert np.array_equal(rid.cpu().numpy(), rr) and np.array_equal(sid.cpu().numpy(), ss)
    # composite invariants: 0 <= acc <= 1 (+eps), rgb in [0,1], white background where nothing is hit
    acc = m.last_aux['acc']
    assert float(acc.min()) >= 0 and float(acc.max()) <= 1 + 1e-5
    assert float(rgb_map.min()) >= 0 and float(rgb_map.max()) <= 1
    # idempotence: same inputs -> same compaction
    rgb2, _, _ = m(torch.from_numpy(rays).cuda(), white_bg=True, is_train=True, N_samples=443)
    assert torch.equal(m.last_aux['samp']['sample_id'], s['sample_id']) and torch.allclose(rgb2, rgb_map, atol=1e-6)
    # empty: rays that miss the box
    miss = torch.tensor([[5., 5., 5., 0., 0., 1.]] * 7, device='cuda')
    rgb3, d3, c3 = m(miss, white_bg=True, is_train=False, N_samples=64)
    assert torch.allclose(rgb3, torch.ones_like(rgb3)) and m.last_aux['samp']['n_valid'] == 0


def test_lazy_counts_match_exact():
    """lazy_counts (device-side counts, capacity-sized buffers, no host sync) gives the same pixels and gradients as the
    exact-size path."""
    from tests import gpu_helpers as G
    g = H.golden('render_train_alpha')
    from ffb200.models.FactorFields import AlphaGridMask
    res = []
    for lazy in (False, True):
        cfg, m = G.build_model(g)
        m.alphaMask = AlphaGridMask('cuda', G.t(g['alpha_aabb']), G.t(g['alpha_volume']))
        m._jitter = lambda n, tr: G.t(g['jitter'])
        m.lazy_counts = lazy
        rgb, depth, coeffs = m(G.t(g['rays']), white_bg=True, is_train=True, N_samples=int(g['N_samples']))
        loss = torch.mean((rgb - G.t(g['target'])) ** 2)
        params = [p for _, p in m.named_parameters()]
        grads = torch.autograd.grad(loss, params, allow_unused=True)
        assert int(m.last_stats['n_valid']) == int(g['n_valid'])
        res.append((rgb, depth, grads, int(m.last_stats['n_app'])))
    assert torch.allclose(res[0][0], res[1][0], atol=1e-7) and torch.allclose(res[0][1], res[1][1], atol=1e-6)
    assert res[0][3] == res[1][3]
    for a, b in zip(res[0][2], res[1][2]):
        assert H.rel_err(G.npy(b), G.npy(a)) < 5e-5
