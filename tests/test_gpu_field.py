"""GPU parity of the field query (K1) and its scatter-add backward (K2) through the C ABI:
 - against the golden vectors produced by the unmodified reference,
 - against the oracle on fresh seeded inputs,
 - size-independent properties at the BASELINE (nerf.yaml) size.
Tolerance: north_star asks for 1e-4 relative in fp32; we assert tighter (see TOL_*)."""
import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL_FWD = 2e-5    # features / coefficients, relative to the largest magnitude
TOL_BWD = 1e-4    # scatter-add gradients (fp32 atomics, order-dependent rounding)


@pytest.mark.parametrize('name', H.field_cases())
def test_field_golden(name):
    from tests import gpu_helpers as G
    g = H.golden('field_' + name)
    cfg, m = G.build_model(g)
    assert list(m.basis_reso) == list(g['fact.basis_reso']) or np.isnan(np.array(m.basis_reso, float)).any()
    assert np.array_equal(G.npy(m.freq_bands), g['fact.freq_bands'])
    assert int(m.n_parameters()) == int(g['fact.n_parameters'])
    x = G.t(g['x'])
    feats, coeff = m.get_coding(x)
    assert H.rel_err(G.npy(feats), g['feats']) < TOL_FWD, name
    assert H.rel_err(G.npy(coeff), g['coeff']) < TOL_FWD, name
    y = m.linear_mat(feats)
    assert H.rel_err(G.npy(y), g['linear_mat_out']) < 1e-4, name
    params = [(n, p) for n, p in m.named_parameters() if n.startswith('coeffs') or n.startswith('basises')]
    if params:
        grads = torch.autograd.grad((feats * G.t(g['G'])).sum(), [p for _, p in params], allow_unused=True)
        for (n, p), gr in zip(params, grads):
            ref = g['grad.' + n]
            got = G.npy(gr) if gr is not None else np.zeros_like(ref)
            assert got.shape == ref.shape
            assert H.rel_err(got, ref) < TOL_BWD, (name, n)


@pytest.mark.parametrize('name', ['nerf_grid_box', 'image', 'nerf_vm', 'nerf_CP', 'image_set', 'nerf_tria', 'sdf'])
def test_field_vs_oracle_random(name):
    """Fresh seeded inputs (incl. points outside the box) at a size the oracle finishes in seconds."""
    from oracle import ff_oracle as O
    from tests import gpu_helpers as G
    g = H.golden('field_' + name)
    cfg, m = G.build_model(g)
    rng = np.random.RandomState(123)
    lo, hi = g['fact.aabb'][0], g['fact.aabb'][1]
    N = 20000
    x = (lo - 0.05 * (hi - lo) + rng.rand(N, lo.size) * 1.1 * (hi - lo)).astype(np.float32)
    if name == 'image':
        x = np.floor(x) + 0.5
    if name == 'image_set':
        x[:, -1] = np.clip(np.floor(x[:, -1]) + 0.5, 0.5, hi[-1] - 0.5)
    Gm = rng.randn(N, g['feats'].shape[1]).astype(np.float32)
    fo = O.FieldOracle(H.oracle_spec(g), H.oracle_params(g))
    f_ref, c_ref = fo.get_coding(x)
    feats, coeff = m.get_coding(G.t(x))
    assert H.rel_err(G.npy(feats), f_ref) < TOL_FWD
    assert H.rel_err(G.npy(coeff), c_ref) < TOL_FWD
    ref = fo.get_coding_bwd(x, Gm)
    params = [(n, p) for n, p in m.named_parameters() if n.startswith('coeffs') or n.startswith('basises')]
    grads = torch.autograd.grad((feats * G.t(Gm)).sum(), [p for _, p in params])
    for (n, p), gr in zip(params, grads):
        kind, i = n.split('.')[:2]
        assert H.rel_err(G.npy(gr), ref[kind][int(i)]) < TOL_BWD, (name, n)


def test_get_coeff_get_basis_split():
    from tests import gpu_helpers as G
    g = H.golden('field_nerf_vm')
    cfg, m = G.build_model(g)
    x = G.t(g['x'])
    feats, coeff = m.get_coding(x)
    c = m.get_coeff(x)
    b = m.get_basis(x)
    assert torch.equal(c, coeff)
    assert H.rel_err(G.npy(b * c), G.npy(feats)) < 1e-6


def test_grid_mapping_all_modes():
    from oracle import ff_oracle as O
    from ffb200.models.FactorFields import grid_mapping
    rng = np.random.RandomState(5)
    aabb = np.array([[-1.2, -0.7, -1.0], [1.3, 0.9, 0.8]], np.float32)
    x = (aabb[0] + rng.rand(5000, 3) * (aabb[1] - aabb[0])).astype(np.float32)
    freq = np.array([1.9, 3.1, 4.3, 5.4, 6.6, 7.8], np.float32)
    for mode in ['sawtooth', 'triangle', 'sinc', 'trigonometric', 'x']:
        ref = O.grid_mapping(x, freq, aabb, mode)
        got = grid_mapping(torch.from_numpy(x).cuda(), torch.from_numpy(freq).cuda(), torch.from_numpy(aabb).cuda(), mode)
        got = got.cpu().numpy()
        assert got.shape == ref.shape
        if mode in ('sawtooth', 'triangle', 'x'):
            assert np.array_equal(got, ref), mode      # pure fp32 arithmetic: bit-exact
        else:
            assert np.abs(got - ref).max() < 2e-6, mode  # sinf/cosf vs libm


def test_full_size_properties():
    """nerf.yaml at its real size (5.3 M parameters, 1 M queries): linearity in the coefficients, weight-sum
    checksum of the scatter-add, empty input."""
    import ffb200
    from ffb200.models.FactorFields import FactorFields
    cfg = ffb200.load_cfg('nerf.yaml')
    cfg.dataset.aabb = [[-1., -1., -1.], [1., 1., 1.]]
    torch.manual_seed(0)
    m = FactorFields(cfg, 'cuda')
    assert m.n_parameters() == 5347600 and m.nSamples == 440
    with torch.no_grad():
        m.coeffs[0].add_(0.1 * torch.randn_like(m.coeffs[0]))
    N = 1 << 20
    x = (torch.rand(N, 3, device='cuda') * 2 - 1)
    feats, coeff = m.get_coding(x)
    assert feats.shape == (N, 18) and torch.isfinite(feats).all()
    # linearity: doubling the coefficient grid doubles coeff and feats exactly (power of two)
    with torch.no_grad():
        m.coeffs[0].mul_(2.0)
    f2, c2 = m.get_coding(x)
    assert torch.equal(c2, coeff * 2) and torch.equal(f2, feats * 2)
    # checksum: trilinear weights of an in-box query sum to 1 (border padding), so
    # sum(grad_coeff[:, c]) == sum_n g[n, c] * basis[n, c]
    gmat = torch.randn(N, 18, device='cuda')
    gc, = torch.autograd.grad((f2 * gmat).sum(), [m.coeffs[0]])
    basis = m.get_basis(x)
    lhs = gc.double().sum(dim=(0, 2, 3, 4))
    rhs = (gmat.double() * basis.double()).sum(0)
    assert torch.allclose(lhs, rhs, rtol=2e-4, atol=1e-2)
    # empty input
    fe, ce = m.get_coding(torch.zeros(0, 3, device='cuda'))
    assert fe.shape == (0, 18)


def test_cpu_tensor_rejected():
    from tests import gpu_helpers as G
    g = H.golden('field_nerf_grid')
    cfg, m = G.build_model(g)
    with pytest.raises(RuntimeError):
        m.get_coding(torch.zeros(4, 3))


@pytest.mark.parametrize('name', ['nerf_grid_box', 'nerf_nearest', 'nerf_tria', 'nerf_sinc', 'nerf_SL', 'sdf', 'image', 'image_bilinear', 'image_set'])
def test_fast_path_matches_generic(name):
    """The specialised grid x grid kernels (field_fast.cu) against the descriptor-driven generic kernels, same inputs."""
    import ctypes as C
    from ffb200 import native as nv
    from tests import gpu_helpers as G
    g = H.golden('field_' + name)
    cfg, m = G.build_model(g)
    plan = m._plan('coding')
    lib = nv.lib()
    assert lib.ffb_field_fast_eligible(plan.handle) == 1, name
    rng = np.random.RandomState(7)
    lo, hi = g['fact.aabb'][0], g['fact.aabb'][1]
    N = 50001
    x = (lo - 0.02 * (hi - lo) + rng.rand(N, lo.size) * 1.04 * (hi - lo)).astype(np.float32)
    x[:4] = g['x'][:4]
    if name == 'image':
        x = np.floor(x) + 0.5
    xd = G.t(x)
    W = plan.width
    out = {k: torch.empty(N, W, device='cuda') for k in ('fg', 'cg', 'ff', 'cf')}
    nv.check(lib.ffb_field_generic_fwd(plan.handle, nv.ptr(xd), C.c_int64(N), None, nv.ptr(out['fg']), nv.ptr(out['cg']), None, nv.stream()))
    nv.check(lib.ffb_field_fast_fwd(plan.handle, nv.ptr(xd), C.c_int64(N), None, nv.ptr(out['ff']), nv.ptr(out['cf']), nv.stream()))
    assert H.rel_err(G.npy(out['ff']), G.npy(out['fg'])) < 2e-6
    assert H.rel_err(G.npy(out['cf']), G.npy(out['cg'])) < 2e-6
    gf = torch.randn(N, W, device='cuda')
    gc = torch.randn(N, W, device='cuda')
    res = []
    for fn in (lib.ffb_field_generic_bwd, lib.ffb_field_fast_bwd):
        grads = [torch.zeros_like(t) for t in plan.tensors]
        arr = (C.c_void_p * nv.MAX_OPS)(*[gr.data_ptr() for gr in grads])
        nv.check(fn(plan.handle, nv.ptr(xd), C.c_int64(N), None, nv.ptr(gf), nv.ptr(gc), arr, nv.stream()))
        res.append(grads)
    for a, b in zip(*res):
        assert H.rel_err(G.npy(b), G.npy(a)) < 5e-5


@pytest.mark.parametrize('name', ['nerf_grid_box', 'nerf_vm', 'nerf_CP'])
def test_deterministic_scatter_mode(name):
    """ffb_set_tuning("field_deterministic", 1): the scatter-add runs serially in query order — two runs give bit-identical
    gradients (the atomic path only agrees to rounding), and they equal the reference's within the usual tolerance."""
    from ffb200 import native as nv
    from tests import gpu_helpers as G
    g = H.golden('field_' + name)
    cfg, m = G.build_model(g)
    rng = np.random.RandomState(17)
    lo, hi = g['fact.aabb'][0], g['fact.aabb'][1]
    x = G.t((lo + rng.rand(6000, lo.size) * (hi - lo)).astype(np.float32))
    x[:g['x'].shape[0]] = G.t(g['x'])
    Gm = G.t(rng.randn(6000, g['feats'].shape[1]).astype(np.float32))
    Gm[:g['x'].shape[0]] = G.t(g['G'])
    params = [(n, p) for n, p in m.named_parameters() if n.startswith('coeffs') or n.startswith('basises')]
    runs = []
    nv.check(nv.lib().ffb_set_tuning(b'field_deterministic', 1))
    try:
        for _ in range(2):
            feats, _ = m.get_coding(x)
            runs.append(torch.autograd.grad((feats * Gm).sum(), [p for _, p in params]))
    finally:
        nv.check(nv.lib().ffb_set_tuning(b'field_deterministic', 0))
    feats, _ = m.get_coding(x)
    atomic = torch.autograd.grad((feats * Gm).sum(), [p for _, p in params])
    for a, b, c in zip(runs[0], runs[1], atomic):
        assert torch.equal(a, b)                                    # bit-reproducible
        assert H.rel_err(G.npy(c), G.npy(a)) < TOL_BWD
    # the golden points alone reproduce the reference's gradients
    nv.check(nv.lib().ffb_set_tuning(b'field_deterministic', 1))
    try:
        n0 = g['x'].shape[0]
        f0, _ = m.get_coding(x[:n0])
        gr = torch.autograd.grad((f0 * Gm[:n0]).sum(), [p for _, p in params])
    finally:
        nv.check(nv.lib().ffb_set_tuning(b'field_deterministic', 0))
    for (n, p), a in zip(params, gr):
        assert H.rel_err(G.npy(a), g['grad.' + n]) < TOL_BWD, n


@pytest.mark.parametrize('name', ['nerf_grid_box', 'nerf_vm', 'nerf_CP', 'image', 'image_set'])
@pytest.mark.parametrize('n_live', [0, 1, 31, 33, 701])
def test_specialised_kernels_ragged_batches_and_device_counts(name, n_live):
    """Ragged sizes (empty, one query, one short of / one past a warp tile, a partial last tile) and the TrainStep calling
    convention — capacity-sized buffers with the live count in device memory (n_dev) — through the product dispatch
    (fast / wide-row / vm / CP kernels), against the generic kernels on exactly n_live rows.  Rows past the live count must
    not contribute to any gradient."""
    import ctypes as C
    from ffb200 import native as nv
    from tests import gpu_helpers as G
    g = H.golden('field_' + name)
    cfg, m = G.build_model(g)
    lib = nv.lib()
    plan = m._plan('coding')
    rng = np.random.RandomState(3 + n_live)
    lo, hi = g['fact.aabb'][0], g['fact.aabb'][1]
    cap = 777
    x = (lo + rng.rand(cap, lo.size) * (hi - lo)).astype(np.float32)
    if name == 'image':
        x = np.floor(x) + 0.5
    if name == 'image_set':
        x[:, -1] = np.clip(np.floor(x[:, -1]) + 0.5, 0.5, hi[-1] - 0.5)
    W = plan.width
    xd, gf = G.t(x), G.t(rng.randn(cap, W).astype(np.float32))
    n_dev = torch.tensor([n_live], dtype=torch.int32, device='cuda')

    def run(train_pair, n, nd):
        f = torch.full((cap, W), 7.0, device='cuda')
        c = torch.full((cap, W), 7.0, device='cuda')
        basis = torch.zeros((cap + 31) // 32 * 32, W, device='cuda')
        grads = [torch.zeros_like(t) for t in plan.tensors]
        arr = (C.c_void_p * nv.MAX_OPS)(*[t.data_ptr() for t in grads])
        if train_pair:
            nv.check(lib.ffb_field_query_fwd_train(plan.handle, nv.ptr(xd), C.c_int64(n), nv.i32p(nd), nv.ptr(f), nv.ptr(c), nv.ptr(basis), nv.stream()))
            nv.check(lib.ffb_field_query_bwd_saved(plan.handle, nv.ptr(xd), C.c_int64(n), nv.i32p(nd), nv.ptr(gf), None, nv.ptr(c), nv.ptr(basis), arr,
                                                   nv.stream()))
        else:
            nv.check(lib.ffb_field_generic_fwd(plan.handle, nv.ptr(xd), C.c_int64(n), nv.i32p(nd), nv.ptr(f), nv.ptr(c), None, nv.stream()))
            nv.check(lib.ffb_field_generic_bwd(plan.handle, nv.ptr(xd), C.c_int64(n), nv.i32p(nd), nv.ptr(gf), None, arr, nv.stream()))
        return f, c, grads

    fp, cp, gp = run(True, cap, n_dev)              # capacity-sized launch, live count on the device
    fg, cg, gg = run(False, n_live, None)           # exactly n_live rows through the generic kernels
    if n_live:
        assert H.rel_err(G.npy(fp[:n_live]), G.npy(fg[:n_live])) < 2e-6
        assert H.rel_err(G.npy(cp[:n_live]), G.npy(cg[:n_live])) < 2e-6
    assert bool((fp[n_live:] == 7.0).all()) and bool((cp[n_live:] == 7.0).all())      # nothing written past the live count
    for a, b, t in zip(gg, gp, plan.tensors):
        scale = max(float(a.abs().max()), 1e-30)
        assert float((a - b).abs().max()) <= 5e-5 * scale, (name, n_live, tuple(t.shape))
