"""Pins oracle/ff_oracle.py against the golden vectors that tests/golden/make_golden.py produced by
running the unmodified reference (CPU fp32).  CPU-only."""
import numpy as np
import pytest

from oracle import ff_oracle as O
from tests import helpers as H

TOL = 2e-6   # oracle vs reference, relative to the largest magnitude (both fp32 on CPU)


@pytest.mark.parametrize('name', H.field_cases())
def test_field_forward_and_grads(name):
    g = H.golden('field_' + name)
    fo = O.FieldOracle(H.oracle_spec(g), H.oracle_params(g))
    feats, coeff = fo.get_coding(g['x'])
    assert feats.shape == g['feats'].shape
    assert H.rel_err(feats, g['feats']) < TOL, name
    assert H.rel_err(coeff, g['coeff']) < TOL, name
    y = O.mlp_forward(H._layers(g, 'param.linear_mat'), feats)
    assert H.rel_err(y, g['linear_mat_out']) < 2e-5
    if any(k.startswith('grad.') for k in g):
        grads = fo.get_coding_bwd(g['x'], g['G'])
        for kind in ('coeffs', 'basises'):
            for i, gr in enumerate(grads[kind]):
                key = f'grad.{kind}.{i}'
                if key in g:
                    assert gr.shape == g[key].shape
                    assert H.rel_err(gr, g[key]) < 1e-5, (name, key)
                elif isinstance(gr, list):     # 'mlp' factor types: per-layer (gW, gb)
                    for l, (gW, gb) in enumerate(gr):
                        assert H.rel_err(gW, g[f'{key}.backbone.{l}.weight']) < 2e-5, (name, key, l)
                        if gb is not None:
                            assert H.rel_err(gb, g[f'{key}.backbone.{l}.bias']) < 2e-5, (name, key, l)


def test_mlps():
    g = H.golden('mlp')
    for tag in ('lm_nerf', 'lm_sdf', 'mlpC', 'deep'):
        i, o, L, hdim, pe = g[f'{tag}.cfg']
        layers = H._layers(g, f'{tag}.param')
        y, cache = O.mlp_forward(layers, g[f'{tag}.x'], pe=int(pe), want_cache=True)
        assert H.rel_err(y, g[f'{tag}.y']) < 1e-5
        gx, grads = O.mlp_backward(layers, cache, g[f'{tag}.G'])
        assert H.rel_err(gx, g[f'{tag}.gx']) < 1e-5
        for l, (gW, gb) in enumerate(grads):
            assert H.rel_err(gW, g[f'{tag}.grad.backbone.{l}.weight']) < 1e-5
            if gb is not None:
                assert H.rel_err(gb, g[f'{tag}.grad.backbone.{l}.bias']) < 1e-5
    layers = H._layers(g, 'rm.param')
    y, cache = O.render_mlp_forward(layers, g['rm.vd'], g['rm.feat'], want_cache=True)
    assert H.rel_err(y, g['rm.y']) < 1e-5
    gf, grads = O.render_mlp_backward(layers, cache, g['rm.G'])
    assert H.rel_err(gf, g['rm.gfeat']) < 1e-5
    for l, (gW, gb) in enumerate(grads):
        assert H.rel_err(gW, g[f'rm.grad.mlp.{l}.weight']) < 1e-5


def test_sampler_bit_exact():
    g = H.golden('sampler_nerf')
    pts, z, inner = O.sample_point(g['aabb'], g['stepSize'], g['rays'][:, :3], g['rays'][:, 3:], 443, g['jitter'])
    assert np.array_equal(np.packbits(inner), g['inner_train'])
    assert np.array_equal(inner.sum(-1), g['counts_train'])
    assert np.array_equal(z[:, 0], g['z_train_first']) and np.array_equal(z[:, -1], g['z_train_last'])
    assert np.allclose(pts.astype(np.float64).sum((0, 1)), g['pts_train_sum'], rtol=1e-9)
    pts, z, inner = O.sample_point(g['aabb'], g['stepSize'], g['rays'][:, :3], g['rays'][:, 3:], int(g['nSamples']), None)
    assert np.array_equal(np.packbits(inner), g['inner_eval'])
    assert z[0, -1] == g['z_eval_last']


def _render_oracle(g):
    spec = H.oracle_spec(g)
    fo = O.FieldOracle(spec, H.oracle_params(g))
    ov = H.load_ref_cfg(g)[1]
    rspec = dict(aabb=g['fact.aabb'], stepSize=g['fact.stepSize'], distance_scale=25.0, density_shift=-10.0,
                 fea2denseAct=ov.get('renderer.fea2denseAct', 'softplus'), rayMarch_weight_thres=1e-3, view_pe=6, fea_pe=2)
    if 'near_far' in g:
        rspec['near_far'] = g['near_far'].tolist()
    if 'bg_len' in g:
        rspec['bg_len'] = float(g['bg_len'])
    mlps = dict(linear_mat=H._layers(g, 'param.linear_mat'), renderModule=H._layers(g, 'param.renderModule'))
    alpha = dict(volume=g['alpha_volume'], aabb=g['alpha_aabb']) if 'alpha_volume' in g else None
    return O.RenderOracle(fo, rspec, mlps, alpha)


RENDER_CASES = ['train', 'train_alpha', 'eval_alpha', 'ndc_train', 'ndc_eval_alpha', 'unbound_train', 'unbound_eval_alpha']


@pytest.mark.parametrize('name', RENDER_CASES)
def test_render_forward_backward(name):
    g = H.golden('render_' + name)
    ro = _render_oracle(g)
    jitter = g['jitter'] if bool(g['is_train']) else None
    mode = str(g['mode']) if 'mode' in g else 'bounded'
    out = ro.forward(g['rays'], int(g['N_samples']), jitter, white_bg=True, want_cache=True, mode=mode)
    if 'pts_sum' in g:
        assert np.allclose(out['pts'].astype(np.float64).sum((0, 1)), g['pts_sum'], rtol=1e-9, atol=1e-9)
    # bit-exact decisions
    assert np.array_equal(np.packbits(out['ray_valid']), g['ray_valid'])
    assert int(out['ray_valid'].sum()) == int(g['n_valid'])
    assert np.array_equal(out['z'], g['z'])
    assert H.rel_err(out['sigma'], g['sigma']) < 2e-5
    assert H.rel_err(out['weight'], g['weight']) < 2e-5
    app_ref = np.unpackbits(g['app_mask'])[:out['app_mask'].size].reshape(out['app_mask'].shape).astype(bool)
    band = np.abs(g['weight'] - 1e-3) < 1e-7
    assert np.array_equal(out['app_mask'][~band], app_ref[~band])
    assert H.rel_err(out['rgb_map'], g['rgb_map']) < 1e-5
    assert H.rel_err(out['depth_map'], g['depth_map']) < 1e-5
    assert H.rel_err(out['coeffs'], g['coeffs']) < 1e-6
    loss, g_rgb = O.mse_loss_and_grad(out['rgb_map'], g['target'])
    assert abs(float(loss) - float(g['loss'])) < 1e-6
    grads = ro.backward(out['cache'], g_rgb)
    for kind in ('coeffs', 'basises'):
        for i, gr in enumerate(grads[kind]):
            assert H.rel_err(gr, g[f'grad.{kind}.{i}']) < 2e-4, (kind, i)
    for l, (gW, gb) in enumerate(grads['linear_mat']):
        assert H.rel_err(gW, g[f'grad.linear_mat.backbone.{l}.weight']) < 2e-4
        if gb is not None:
            assert H.rel_err(gb, g[f'grad.linear_mat.backbone.{l}.bias']) < 2e-4
    for l, (gW, gb) in enumerate(grads['renderModule']):
        assert H.rel_err(gW, g[f'grad.renderModule.mlp.{l}.weight']) < 2e-4
        if gb is not None:
            assert H.rel_err(gb, g[f'grad.renderModule.mlp.{l}.bias']) < 2e-4


def test_render_nerf_scale():
    """The nerf.yaml-scale render case (1024 rays x 443 samples, 248 374 valid samples; recorded compactly): the oracle
    reproduces the reference's masks bit-exactly and its weights / pixels / loss / factor gradients."""
    g = H.golden('render_train_big')
    ro = _render_oracle(g)
    out = ro.forward(g['rays'], int(g['N_samples']), g['jitter'], white_bg=True, want_cache=True, mode='bounded')
    assert np.array_equal(np.packbits(out['ray_valid']), g['ray_valid'])
    assert int(out['ray_valid'].sum()) == int(g['n_valid']) > 151552
    assert np.array_equal(out['z'][:, 0], g['z_first']) and np.array_equal(out['z'][:, -1], g['z_last'])
    assert np.allclose(out['z'].astype(np.float64).sum(1), g['z_rowsum'], rtol=1e-12)
    assert H.rel_err(out['weight'][out['ray_valid']], g['weight_valid']) < 2e-5
    assert H.rel_err(out['rgb_map'], g['rgb_map']) < 1e-5
    assert H.rel_err(out['depth_map'], g['depth_map']) < 1e-5
    assert H.rel_err(out['coeffs'][::int(g['coeffs_stride'])], g['coeffs_rows']) < 1e-6
    loss, g_rgb = O.mse_loss_and_grad(out['rgb_map'], g['target'])
    assert abs(float(loss) - float(g['loss'])) < 1e-6
    grads = ro.backward(out['cache'], g_rgb)
    for kind in ('coeffs', 'basises'):
        for i, gr in enumerate(grads[kind]):
            assert H.rel_err(gr, g[f'grad.{kind}.{i}']) < 2e-4, (kind, i)
    for l, (gW, gb) in enumerate(grads['linear_mat']):
        assert H.rel_err(gW, g[f'grad.linear_mat.backbone.{l}.weight']) < 2e-4
    for l, (gW, gb) in enumerate(grads['renderModule']):
        assert H.rel_err(gW, g[f'grad.renderModule.mlp.{l}.weight']) < 2e-4


def test_alpha_mask_maintenance():
    """Oracle restatement of compute_alpha / getDenseAlpha (no jitter) / filtering_rays against the vectors recorded from the
    reference (tests/golden/maintenance.npz)."""
    g = H.golden('maintenance')
    ro = _render_oracle(g)
    assert H.rel_err(O.compute_alpha(ro, g['ca_xyz'], 0.2), g['ca_alpha']) < 2e-5
    alpha, dense = O.dense_alpha(ro, g['fact.aabb'], [20, 18, 22])
    assert np.allclose(dense.astype(np.float64).sum((0, 1, 2)), g['dense_xyz_sum'], rtol=1e-7)
    assert H.rel_err(alpha, g['dense_alpha']) < 2e-5
    ro.alpha = dict(volume=g['mask_volume'], aabb=g['mask_aabb'])
    # alpha = 1 - exp(-x) carries the absolute rounding of exp() near 1 (an ulp of 1.0 = 6e-8), and every point the mask keeps
    # here has a small alpha: absolute tolerance
    assert np.abs(O.compute_alpha(ro, g['ca_xyz'], 0.2) - g['ca_alpha_masked']).max() < 3e-7
    keep = O.filter_rays_mask(ro, g['f_rays'], 64)
    assert np.array_equal(g['f_rays'][keep], g['f_kept_rays']) and np.array_equal(g['f_rgbs'][keep], g['f_kept_rgbs'])
    keep_bbox = O.filter_rays_mask(ro, g['f_rays'], 64, bbox_only=True)
    assert np.array_equal(g['f_rays'][keep_bbox], g['f_kept_rays_bbox'])
    # update_renderParams after the upsample (:693-699)
    step, n = O.update_render_params(g['fact.aabb'], g['up_gridSize'], 0.5)
    assert step == np.float32(g['up_stepSize']) and n == int(g['up_nSamples'])


def test_module_surface_functions():
    """positional_encoding (:74-79), raw2alpha (:82-88) and basis2density (:639-643) of the oracle vs the reference's outputs."""
    g = H.golden('api')
    assert H.rel_err(O.positional_encoding(g['pe_x'], 4), g['pe_y']) < 1e-6
    a, w, bg = O.raw2alpha(g['r2a_sigma'], g['r2a_dist'])
    assert H.rel_err(a, g['r2a_alpha']) < 1e-6 and H.rel_err(w, g['r2a_weight']) < 1e-6 and H.rel_err(bg, g['r2a_bg']) < 1e-6
    ro = O.RenderOracle(None, dict(density_shift=-10.0, fea2denseAct='softplus'), None)
    assert H.rel_err(ro.basis2density(g['b2d_f']), g['b2d_softplus']) < 1e-6
    ro.r['fea2denseAct'] = 'relu'
    assert H.rel_err(ro.basis2density(g['b2d_f']), g['b2d_relu']) < 1e-6


def test_torch_port():
    """oracle/torch_port.py (the CPU-baseline restatement with torch CPU operators) against the reference's golden
    render vectors: same masks, rgb, loss."""
    import torch
    from oracle.torch_port import TorchPort
    g = H.golden('render_train')
    state = {k[len('param.'):]: v for k, v in g.items() if k.startswith('param.')}
    tp = TorchPort(state, g['fact.aabb'], g['fact.freq_bands'], g['fact.stepSize'],
                   dict(density_shift=-10.0, distance_scale=25.0, rayMarch_weight_thres=1e-3, view_pe=6, fea_pe=2))
    rgb, depth, valid, weight = tp.forward(torch.from_numpy(g['rays']), int(g['N_samples']), torch.from_numpy(g['jitter']))
    assert np.array_equal(np.packbits(valid.numpy()), g['ray_valid'])
    assert H.rel_err(rgb.detach().numpy(), g['rgb_map']) < 1e-6
    assert H.rel_err(weight.detach().numpy(), g['weight']) < 1e-6
    loss = torch.mean((rgb - torch.from_numpy(g['target'])) ** 2)
    assert abs(float(loss.detach()) - float(g['loss'])) < 1e-7
    gc, = torch.autograd.grad(loss, [tp.p['coeffs.0']])
    assert H.rel_err(gc.numpy(), g['grad.coeffs.0']) < 1e-5


@pytest.mark.parametrize('name,coef_mode,basis_mode', [('image', 'nearest', 'nearest'), ('image_bilinear', 'bilinear', 'bilinear'),
                                                       ('sdf', 'bilinear', 'bilinear'), ('image_set', 'bilinear', 'bilinear')])
def test_regress_port(name, coef_mode, basis_mode):
    """oracle/torch_port.RegressPort (reference side of the regression-driver parity tests and the CPU baseline of the
    image / sdf bench workloads) against the reference's golden field vectors: features, coefficients, MLP output and
    the field gradients of sum(feats * G)."""
    import torch
    from oracle.torch_port import RegressPort
    g = H.golden('field_' + name)
    state = {k[len('param.'):]: v for k, v in g.items() if k.startswith('param.')}
    rp = RegressPort(state, g['fact.aabb'], g['fact.freq_bands'], int(g['fact.in_dim']), coef_mode, basis_mode)
    x = torch.from_numpy(g['x'])
    feats, coeff = rp.get_coding(x)
    assert H.rel_err(feats.detach().numpy(), g['feats']) < 1e-6 and H.rel_err(coeff.detach().numpy(), g['coeff']) < 1e-6
    assert H.rel_err(rp.linear_mat(feats).detach().numpy(), g['linear_mat_out']) < 1e-6
    keys = [k for k in rp.p if k.startswith(('coeffs', 'basises'))]
    grads = torch.autograd.grad((feats * torch.from_numpy(g['G'])).sum(), [rp.p[k] for k in keys])
    for k, gr in zip(keys, grads):
        assert H.rel_err(gr.numpy(), g['grad.' + k]) < 1e-5, k


def test_bench_workload_matches_reference_shapes():
    """bench_workload's hard-coded nerf.yaml facts equal what the product's host logic derives from configs/nerf.yaml."""
    import bench_workload as W
    import ffb200
    from ffb200.models.FactorFields import field_shapes
    cfg = ffb200.load_cfg('nerf.yaml')
    sh = field_shapes(cfg, W.AABB)
    assert list(sh['basis_reso']) == W.BASIS_RESO and [int(v) for v in sh['coeff_reso']] == [W.COEFF_RESO] * 3
    assert np.array_equal(sh['freq_bands'].numpy(), np.array(W.FREQ_BANDS, np.float32))
    sd = W.make_state()
    assert sum(v.size for v in sd.values()) == 5347600
