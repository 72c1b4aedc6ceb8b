"""Element-wise parity of the kernels the BENCH line times, at the sizes it times them.

field_fast.cu picks its instantiation by batch size: up to 151 552 queries (6 levels) the level-parallel kernels, above
that one thread per query with the blocked saved basis row (`fast_fwd_kernel<.., LPAR=0>`) and the run-aggregated scatter
(`fast_bwd_saved_agg_kernel`).  The small golden cases only reach the former, so this file drives the latter:

 * >= 1 M ray-ordered queries at the real nerf.yaml / sdf.yaml shapes and full image.yaml / image_set.yaml shapes through
   `ffb_field_query_fwd_train` / `ffb_field_query_bwd_saved`, element-wise against the descriptor-driven generic kernels
   (features, coefficient row, saved basis row, EVERY factor gradient) and against the oracle on a >= 200 k subset;
 * all reference-generated field cases again with `field_level_parallel = 0` (large-batch instantiations on reference vectors);
 * the nerf.yaml-scale render case (1024 rays x 443 samples, 248 k valid samples) recorded from the unmodified reference.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL_FWD = 2e-5
TOL_BWD = 1e-4     # scatter-add gradients (fp32 atomics: order-dependent rounding), relative to the largest magnitude


def _ray_ordered_points(n_rays, per_ray, lo, hi, seed, step):
    """Consecutive rows = consecutive samples of a ray (the order the sampler produces, which the run-aggregated scatter
    relies on), rays criss-crossing the box; ~2 % of the points fall outside it (border / zero padding paths)."""
    rng = np.random.RandomState(seed)
    d = lo.size
    span = hi - lo
    o = lo - 0.01 * span + rng.rand(n_rays, d) * 1.02 * span
    v = rng.randn(n_rays, d)
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    t = (np.arange(per_ray) - per_ray / 2)[None, :, None] * step
    p = o[:, None, :] + v[:, None, :] * t
    # fold back into a slightly enlarged box so the points stay (mostly) inside
    p = lo - 0.01 * span + np.abs(((p - lo + 0.01 * span) % (2.04 * span)) - 1.02 * span)
    return p.reshape(-1, d).astype(np.float32)


def _full_model(cfgname, aabb, seed=0, overrides=()):
    import ffb200
    from ffb200.models.FactorFields import FactorFields
    cfg = ffb200.load_cfg(cfgname, list(overrides))
    cfg.dataset.aabb = aabb
    torch.manual_seed(seed)
    m = FactorFields(cfg, 'cuda')
    with torch.no_grad():
        for p in list(m.coeffs) + list(m.basises):
            p.add_(0.3 * torch.randn_like(p))
    return cfg, m


def _oracle_for(m, cfg):
    from oracle import ff_oracle as O
    spec = dict(mode=cfg.defaults.mode, in_dim=int(m.in_dim), aabb=m.aabb.cpu().numpy(), freq_bands=m.freq_bands.cpu().numpy(),
                basis_dims=list(m.basis_dims), coeff_type=cfg.model.coeff_type, basis_type=cfg.model.basis_type,
                basis_mapping=cfg.model.basis_mapping, coef_mode=cfg.model.coef_mode, basis_mode=cfg.model.basis_mode)
    params = dict(coeffs=[p.detach().cpu().contiguous().numpy() for p in m.coeffs],
                  basises=[p.detach().cpu().contiguous().numpy() for p in m.basises])
    return O.FieldOracle(spec, params)


def _run_fast_and_generic(m, x, gf, gc):
    """-> dict of outputs of the product entry points (ffb_field_query_fwd_train / _bwd_saved: the kernels the bench
    times) and of the generic kernels on the same inputs."""
    from ffb200 import native as nv
    lib = nv.lib()
    plan = m._plan('coding')
    assert lib.ffb_field_fast_eligible(plan.handle) == 1
    N, W = x.shape[0], plan.width
    out = {k: torch.empty(N, W, device='cuda') for k in ('ff', 'cf', 'fg', 'cg')}
    basis = torch.empty((N + 31) // 32 * 32, W, device='cuda')
    nv.check(lib.ffb_field_query_fwd_train(plan.handle, nv.ptr(x), C.c_int64(N), None, nv.ptr(out['ff']), nv.ptr(out['cf']), nv.ptr(basis), nv.stream()))
    nv.check(lib.ffb_field_generic_fwd(plan.handle, nv.ptr(x), C.c_int64(N), None, nv.ptr(out['fg']), nv.ptr(out['cg']), None, nv.stream()))
    # un-block the saved basis row: element (i, c) at (i / 32) * 32 W + c * 32 + i % 32   (field_fast.cu: blk_idx)
    # (rows of >= 64 channels — image.yaml / image_set.yaml — are saved row-major by the column-parallel kernels)
    layout = lib.ffb_field_saved_basis_layout(plan.handle, C.c_int64(N))
    assert layout in (0, 1)
    out['basis'] = basis[:N] if layout == 1 else basis.view(-1, W, 32).permute(0, 2, 1).reshape(-1, W)[:N]
    g_fast = [torch.zeros_like(t) for t in plan.tensors]
    arr = (C.c_void_p * nv.MAX_OPS)(*[g.data_ptr() for g in g_fast])
    nv.check(lib.ffb_field_query_bwd_saved(plan.handle, nv.ptr(x), C.c_int64(N), None, nv.ptr(gf), nv.ptr(gc, allow_none=True),
                                           nv.ptr(out['cf']), nv.ptr(basis), arr, nv.stream()))
    g_gen = [torch.zeros_like(t) for t in plan.tensors]
    arr = (C.c_void_p * nv.MAX_OPS)(*[g.data_ptr() for g in g_gen])
    nv.check(lib.ffb_field_generic_bwd(plan.handle, nv.ptr(x), C.c_int64(N), None, nv.ptr(gf), nv.ptr(gc, allow_none=True), arr, nv.stream()))
    out['g_fast'], out['g_gen'] = g_fast, g_gen
    return out


FULL = {
    # name: (yaml, aabb, n_rays, per_ray, step (aabb units), subset rays for the oracle, with g_coeff)
    'nerf': ('nerf.yaml', [[-1., -1., -1.], [1., 1., 1.]], 4096, 260, 2.0 / 127 * 0.5, 820, False),
    'sdf': ('sdf.yaml', [[0., 0., 0.], [640., 640., 640.]], 4096, 256, 1.7, 800, True),
    'image': ('image.yaml', [[0., 0.], [1024., 1024.]], 2048, 512, 1.0, 400, False),
    'image_set': ('image_set.yaml', [[0, 0, 0], [256, 256, 40]], 2048, 512, 0.7, 400, False),
}


@pytest.mark.parametrize('name', list(FULL))
def test_timed_kernels_full_size_elementwise(name):
    from tests import gpu_helpers as G
    from ffb200 import native as nv
    yaml_, aabb, n_rays, per_ray, step, sub_rays, with_gc = FULL[name]
    cfg, m = _full_model(yaml_, aabb, overrides=['model.with_dropout=false'] if name == 'image_set' else ())
    if name == 'nerf':
        assert m.n_parameters() == 5347600 and list(m.coeffs[0].shape) == [1, 18, 32, 32, 32]
    if name == 'image':
        assert list(m.coeffs[0].shape) == [1, 144, 63, 63] and cfg.model.coef_mode == 'nearest'
    lo, hi = (np.array(a, np.float64) for a in aabb)
    x = _ray_ordered_points(n_rays, per_ray, lo, hi, 11, step)
    if name == 'image':
        x = np.floor(x) + 0.5
    if name == 'image_set':
        x[:, 2] = np.clip(np.floor(x[:, 2]), 0, hi[2] - 1) + 0.5
    N = x.shape[0]
    assert N >= (1 << 20)
    W = sum(m.basis_dims)
    assert N * len(m.basis_dims) > 148 * 2048 * 3          # above LPAR_MAX_ITEMS: the large-batch instantiations run
    rng = np.random.RandomState(5)
    xd = G.t(x)
    gf = G.t(rng.randn(N, W).astype(np.float32))
    gc = G.t(rng.randn(N, W).astype(np.float32)) if with_gc else None
    for lpar in (1, 0):                                     # 1: product defaults;  0: the same through the other dispatch branch
        nv.check(nv.lib().ffb_set_tuning(b'field_level_parallel', lpar))
        try:
            r = _run_fast_and_generic(m, xd, gf, gc)
        finally:
            nv.check(nv.lib().ffb_set_tuning(b'field_level_parallel', 1))
        assert H.rel_err(G.npy(r['ff']), G.npy(r['fg'])) < 2e-6, name
        assert H.rel_err(G.npy(r['cf']), G.npy(r['cg'])) < 2e-6, name
        nz = r['cg'].abs() > 1e-3                           # saved basis row == feats / coeff wherever the coefficient is not ~0
        assert float(((r['basis'] * r['cg'] - r['fg']).abs() * nz).max()) < 2e-5 * float(r['fg'].abs().max())
        for a, b, t in zip(r['g_gen'], r['g_fast'], m._plan('coding').tensors):
            assert H.rel_err(G.npy(b), G.npy(a)) < 5e-5, (name, tuple(t.shape))
    # ---- the oracle on a ray-ordered subset that is itself above the threshold
    ns = sub_rays * per_ray
    assert ns >= 200000 and ns * len(m.basis_dims) > 148 * 2048 * 3
    xs, gfs = x[:ns], G.npy(gf[:ns])
    gcs = G.npy(gc[:ns]) if gc is not None else None
    fo = _oracle_for(m, cfg)
    f_ref, c_ref = fo.get_coding(xs)
    rs = _run_fast_and_generic(m, G.t(xs), G.t(gfs), G.t(gcs) if gcs is not None else None)
    assert H.rel_err(G.npy(rs['ff']), f_ref) < TOL_FWD, name
    assert H.rel_err(G.npy(rs['cf']), c_ref) < TOL_FWD, name
    ref = fo.get_coding_bwd(xs, gfs, gcs) if gcs is not None else fo.get_coding_bwd(xs, gfs)
    names = [n for n, _ in m.named_parameters() if n.startswith(('coeffs', 'basises'))]
    for n, got in zip(names, rs['g_fast']):
        kind, i = n.split('.')[:2]
        assert H.rel_err(G.npy(got), ref[kind][int(i)]) < TOL_BWD, (name, n)


@pytest.mark.parametrize('name', H.field_cases())
def test_field_golden_large_batch_instantiations(name):
    """Every reference-generated field case with the level-parallel dispatch switched off, so the reference's own vectors
    reach `fast_fwd_kernel<.., LPAR=0>` and `fast_bwd_saved_agg_kernel` / `fast_bwd_saved_kernel<AGGW>` (cases that are not
    eligible for the fast path run the generic kernels as before)."""
    from ffb200 import native as nv
    from tests import gpu_helpers as G
    g = H.golden('field_' + name)
    cfg, m = G.build_model(g)
    nv.check(nv.lib().ffb_set_tuning(b'field_level_parallel', 0))
    try:
        x = G.t(g['x'])
        feats, coeff = m.get_coding(x)
        assert H.rel_err(G.npy(feats), g['feats']) < TOL_FWD, name
        assert H.rel_err(G.npy(coeff), g['coeff']) < TOL_FWD, name
        params = [(n, p) for n, p in m.named_parameters() if n.startswith('coeffs') or n.startswith('basises')]
        if params:
            grads = torch.autograd.grad((feats * G.t(g['G'])).sum(), [p for _, p in params], allow_unused=True)
            for (n, p), gr in zip(params, grads):
                ref = g['grad.' + n]
                got = G.npy(gr) if gr is not None else np.zeros_like(ref)
                assert H.rel_err(got, ref) < TOL_BWD, (name, n)
    finally:
        nv.check(nv.lib().ffb_set_tuning(b'field_level_parallel', 1))


def _unpack(bits, shape):
    return np.unpackbits(bits)[:int(np.prod(shape))].reshape(shape).astype(bool)


@pytest.mark.parametrize('lazy', [False, True])
def test_render_golden_nerf_scale(lazy):
    """forward() + autograd at the nerf.yaml sampling scale (1024 rays x 443 samples, 248 374 valid samples: above the
    level-parallel threshold, so the field kernels are the instantiations the bench times) against the unmodified
    reference: bit-exact sample indices and counts, weights / rgb / depth / coeffs / loss and every parameter gradient.
    lazy=True is the TrainStep configuration (device-side counts, capacity-sized buffers)."""
    from ffb200.renderer import render_ray
    from tests import gpu_helpers as G
    g = H.golden('render_train_big')
    cfg, m = G.build_model(g)
    assert m.nSamples == int(g['fact.nSamples']) and float(m.stepSize) == float(g['fact.stepSize'])
    S, R = int(g['N_samples']), g['rays'].shape[0]
    m._jitter = lambda n, tr: G.t(g['jitter']) if tr else None
    m.lazy_counts = lazy
    rgb_map, depth_map, coeffs = render_ray(torch.from_numpy(g['rays']), m, chunk=4096, N_samples=S, white_bg=True, is_train=True, device='cuda')
    aux = m.last_aux
    valid_ref = _unpack(g['ray_valid'], (R, S))
    rr, ss = np.nonzero(valid_ref)
    nv_ = int(g['n_valid'])
    assert nv_ > 151552 and int(aux['samp']['n_valid']) == nv_
    assert np.array_equal(G.npy(aux['samp']['ray_id'])[:nv_], rr.astype(np.int32))
    assert np.array_equal(G.npy(aux['samp']['sample_id'])[:nv_], ss.astype(np.int32))
    assert np.array_equal(G.npy(aux['samp']['counts']), valid_ref.sum(-1).astype(np.int32))
    z = G.npy(aux['samp']['z'])[:nv_]
    for col, key in ((0, 'z_first'), (S - 1, 'z_last')):     # interpx of the reference at the first / last sample index, bit-exact
        sel = ss == col                                       # (rays leave the box before the last index: that set may be empty)
        assert np.array_equal(z[sel], g[key][rr[sel]])
    assert (ss == 0).any()
    w_ref = g['weight_valid']
    assert H.rel_err(G.npy(aux['weight'])[:nv_], w_ref) < 1e-4
    app_ref = _unpack(g['app_mask'], (R, S))[valid_ref]
    n_app = int(m.last_stats['n_app'])
    app = np.zeros(nv_, bool)
    app[G.npy(aux['app_idx'])[:n_app]] = True
    band = np.abs(w_ref - 1e-3) < 1e-6
    assert np.array_equal(app[~band], app_ref[~band])
    assert abs(n_app - int(g['n_app'])) <= int(band.sum())
    assert H.rel_err(G.npy(rgb_map), g['rgb_map']) < 1e-4
    assert H.rel_err(G.npy(depth_map), g['depth_map']) < 1e-4
    cf = G.npy(coeffs)[:nv_]
    assert cf.shape == tuple(g['coeffs_shape'])
    assert H.rel_err(cf[::int(g['coeffs_stride'])], g['coeffs_rows']) < TOL_FWD
    assert np.allclose(cf.astype(np.float64).sum(0), g['coeffs_colsum'], rtol=1e-6)
    loss = torch.mean((rgb_map - G.t(g['target'])) ** 2)
    assert abs(float(loss) - float(g['loss'])) < 1e-5
    params = list(m.named_parameters())
    grads = torch.autograd.grad(loss, [p for _, p in params], allow_unused=True)
    for (n, p), gr in zip(params, grads):
        ref = g['grad.' + n]
        got = G.npy(gr) if gr is not None else np.zeros_like(ref)
        assert H.rel_err(got, ref) < 2e-4, n


PRESET_FULL = {
    # BASELINE config 4 at its real shapes: (model overrides, expected specialised path)
    'nerf_cp': (['model.coeff_type=vec', 'model.basis_type=cp', 'model.freq_bands=[1.,1.,1.,1.,1.,1.]', 'model.basis_resos=[512,512,512,512,512,512]',
                 'model.basis_dims=[32,32,32,32,32,32]'], 'lines'),
    'nerf_vm': (['model.coeff_type=vm', 'model.basis_type=vm'], 'planes'),
}


@pytest.mark.parametrize('name', list(PRESET_FULL))
def test_preset_kernels_full_shape(name):
    """The -CP / -vm presets (README_FactorField.md:12-32) at the Tanks&Temples bench shapes (non-cubic box, 1 M ray-ordered
    queries): the specialised kernels the ffb_field_query_* entry points dispatch to, element-wise against the descriptor-driven
    generic kernels (every output, every factor gradient, with and without an upstream coefficient gradient) and against the
    oracle on a 200 k subset."""
    import bench_workload as W
    import ffb200
    from ffb200 import native as nv
    from ffb200.models.FactorFields import FactorFields
    from tests import gpu_helpers as G
    ov, kind = PRESET_FULL[name]
    cfg = ffb200.load_cfg('nerf.yaml', ov)
    cfg.dataset.aabb = W.TNT_AABB
    torch.manual_seed(1)
    m = FactorFields(cfg, 'cuda')
    with torch.no_grad():
        for p in list(m.coeffs) + list(m.basises):
            p.add_(0.3 * torch.randn_like(p))
    lib = nv.lib()
    plan = m._plan('coding')
    if kind == 'lines':
        assert lib.ffb_field_lines_eligible(plan.handle) == 1
    if kind == 'planes':
        assert lib.ffb_field_planes_eligible(plan.handle) == 1
    lo, hi = (np.array(a, np.float64) for a in W.TNT_AABB)
    x = _ray_ordered_points(4096, 256, lo, hi, 21, float(m.stepSize))
    N, Wd = x.shape[0], plan.width
    rng = np.random.RandomState(6)
    xd = G.t(x)
    gf = G.t(rng.randn(N, Wd).astype(np.float32))
    gc = G.t(rng.randn(N, Wd).astype(np.float32))

    def run(fwd, bwd, with_gc):
        f, c = torch.empty(N, Wd, device='cuda'), torch.empty(N, Wd, device='cuda')
        nv.check(fwd(plan.handle, nv.ptr(xd), C.c_int64(N), None, nv.ptr(f), nv.ptr(c), nv.stream()))
        grads = [torch.zeros_like(t) for t in plan.tensors]
        arr = (C.c_void_p * nv.MAX_OPS)(*[g.data_ptr() for g in grads])
        nv.check(bwd(plan.handle, nv.ptr(xd), C.c_int64(N), None, nv.ptr(gf), nv.ptr(gc) if with_gc else None, arr, nv.stream()))
        return f, c, grads

    gen_fwd = lambda h, x_, n, nd, f, c, s: lib.ffb_field_generic_fwd(h, x_, n, nd, f, c, None, s)
    for with_gc in (False, True):
        fp, cp, gp = run(lib.ffb_field_query_fwd, lib.ffb_field_query_bwd, with_gc)          # product dispatch
        fg, cg, gg = run(gen_fwd, lib.ffb_field_generic_bwd, with_gc)
        assert H.rel_err(G.npy(fp), G.npy(fg)) < 2e-6 and H.rel_err(G.npy(cp), G.npy(cg)) < 2e-6
        for a, b, t in zip(gg, gp, plan.tensors):
            assert H.rel_err(G.npy(b), G.npy(a)) < 5e-5, (name, with_gc, tuple(t.shape))
    ns = 800 * 256
    fo = _oracle_for(m, cfg)
    f_ref, c_ref = fo.get_coding(x[:ns])
    feats, coeff = m.get_coding(G.t(x[:ns]))
    assert H.rel_err(G.npy(feats), f_ref) < TOL_FWD and H.rel_err(G.npy(coeff), c_ref) < TOL_FWD
    ref = fo.get_coding_bwd(x[:ns], G.npy(gf[:ns]))
    params = [(n_, p) for n_, p in m.named_parameters() if n_.startswith(('coeffs', 'basises'))]
    grads = torch.autograd.grad((feats * gf[:ns]).sum(), [p for _, p in params])
    for (n_, p), gr in zip(params, grads):
        kind_, i = n_.split('.')[:2]
        assert H.rel_err(G.npy(gr), ref[kind_][int(i)]) < TOL_BWD, (name, n_)
