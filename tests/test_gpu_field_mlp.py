"""The field query fused with linear_mat (field_mlp.cu) through the C ABI, against the separate kernels it replaces
(ffb_field_query_fwd_train + ffb_mlp2_fwd / ffb_mlp2_bwd + ffb_field_query_bwd_saved), against the oracle, and against the
reference-generated vectors."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _unblock(t, n, W):
    return t.view(-1, W, 32).permute(0, 2, 1).reshape(-1, W)[:n]


def _nerf_model(total=None, seed=0):
    import ffb200
    from ffb200.models.FactorFields import FactorFields
    ov = [f'model.total_params={total}', 'model.coeff_reso=8'] if total else []
    cfg = ffb200.load_cfg('nerf.yaml', ov)
    cfg.dataset.aabb = [[-1., -1., -1.], [1., 1., 1.]]
    torch.manual_seed(seed)
    m = FactorFields(cfg, 'cuda')
    with torch.no_grad():
        for p in list(m.coeffs) + list(m.basises):
            p.add_(0.3 * torch.randn_like(p))
        m.linear_mat.backbone[0].weight.mul_(3.0)
    return cfg, m


def _points(n, seed):
    from tests.test_gpu_fullsize import _ray_ordered_points
    lo, hi = np.array([-1., -1., -1.]), np.array([1., 1., 1.])
    per = 256
    x = _ray_ordered_points((n + per - 1) // per, per, lo, hi, seed, 2.0 / 127 * 0.5)[:n]
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _fused_fwd(m, x, want_rows=True):
    from ffb200 import native as nv
    lib = nv.lib()
    plan = m._plan('coding')
    W1, b1, W2 = m.linear_mat.backbone[0].weight, m.linear_mat.backbone[0].bias, m.linear_mat.backbone[1].weight
    K0, Hd, N = W1.shape[1], W1.shape[0], W2.shape[0]
    assert lib.ffb_field_mlp_eligible(plan.handle, K0, Hd, N) == 1
    n, W = x.shape[0], plan.width
    nb = (n + 31) // 32 * 32
    out = dict(y=torch.empty(n, N, device='cuda'), bits=torch.empty(n, 4, device='cuda', dtype=torch.int16),
               cb=torch.empty(nb, W, device='cuda'), bb=torch.empty(nb, W, device='cuda'),
               feats=torch.empty(n, W, device='cuda') if want_rows else None, coeff=torch.empty(n, W, device='cuda') if want_rows else None)
    nv.check(lib.ffb_field_mlp_fwd(plan.handle, nv.ptr(x), C.c_int64(n), None, nv.ptr(W1), nv.ptr(b1), nv.ptr(W2), nv.ptr(out['y']),
                                   nv.ptr(out['bits'], torch.int16), nv.ptr(out['cb']), nv.ptr(out['bb']), nv.ptr(out['feats'], allow_none=True),
                                   nv.ptr(out['coeff'], allow_none=True), K0, Hd, N, nv.stream()))
    return out


def _separate_fwd(m, x):
    from ffb200 import native as nv
    lib = nv.lib()
    plan = m._plan('coding')
    W1, b1, W2 = m.linear_mat.backbone[0].weight, m.linear_mat.backbone[0].bias, m.linear_mat.backbone[1].weight
    n, W = x.shape[0], plan.width
    feats, coeff = torch.empty(n, W, device='cuda'), torch.empty(n, W, device='cuda')
    basis = torch.empty((n + 31) // 32 * 32, W, device='cuda')
    nv.check(lib.ffb_field_query_fwd_train(plan.handle, nv.ptr(x), C.c_int64(n), None, nv.ptr(feats), nv.ptr(coeff), nv.ptr(basis), nv.stream()))
    y = torch.empty(n, W2.shape[0], device='cuda')
    bits = torch.empty(n, 4, device='cuda', dtype=torch.int16)
    nv.check(lib.ffb_mlp2_fwd(nv.ptr(feats), nv.ptr(W1), nv.ptr(b1), nv.ptr(W2), nv.ptr(y), nv.ptr(bits, torch.int16), C.c_int64(n), None,
                              W1.shape[1], W1.shape[0], W2.shape[0], nv.stream()))
    return dict(y=y, bits=bits, feats=feats, coeff=coeff, basis=basis)


@pytest.mark.parametrize('n', [1, 127, 128, 4097, 300001, 1 << 20])
def test_fused_forward_matches_separate_kernels(n):
    """Every output of the fused forward kernel equals the separate kernels': features / coefficient rows bit-exactly (same
    arithmetic), y to fp32 rounding (same split products, different accumulation grouping), ReLU bits wherever the hidden
    value is not within rounding distance of zero."""
    from tests import gpu_helpers as G
    cfg, m = _nerf_model(total=None if n >= 4097 else 60000)
    x = _points(n, 3)
    f, s = _fused_fwd(m, x), _separate_fwd(m, x)
    torch.cuda.synchronize()
    W = m._plan('coding').width
    assert torch.equal(f['feats'], s['feats'])
    assert torch.equal(f['coeff'], s['coeff'])
    assert torch.equal(_unblock(f['cb'], n, W), s['coeff'])
    assert torch.equal(_unblock(f['bb'], n, W), _unblock(s['basis'], n, W))
    assert H.rel_err(G.npy(f['y']), G.npy(s['y'])) < 2e-6
    # fp64 reference of linear_mat on the (identical) feature rows
    W1, b1, W2 = (p.detach().double() for p in (m.linear_mat.backbone[0].weight, m.linear_mat.backbone[0].bias, m.linear_mat.backbone[1].weight))
    hid = s['feats'].double() @ W1.t() + b1
    y64 = torch.relu(hid) @ W2.t()
    assert H.rel_err(G.npy(f['y']), G.npy(y64)) < 2e-6
    bits = (f['bits'].view(torch.uint8).view(n, 8).to(torch.int32))
    dec = torch.stack([(bits[:, k // 8] >> (k % 8)) & 1 for k in range(64)], 1).bool()
    clear = hid.abs() > 1e-5 * hid.abs().max()
    assert torch.equal(dec[clear], (hid > 0)[clear])


def test_fused_forward_golden_and_oracle():
    """Reference vectors (301 points, nerf_grid case) and the oracle on 200 k fresh points through the fused kernel."""
    from oracle import ff_oracle as O
    from tests import gpu_helpers as G
    g = H.golden('field_nerf_grid_box')
    cfg, m = G.build_model(g)
    f = _fused_fwd(m, G.t(g['x']))
    assert H.rel_err(G.npy(f['feats']), g['feats']) < 2e-5
    assert H.rel_err(G.npy(f['coeff']), g['coeff']) < 2e-5
    assert H.rel_err(G.npy(f['y']), g['linear_mat_out']) < 1e-4
    rng = np.random.RandomState(9)
    lo, hi = g['fact.aabb'][0], g['fact.aabb'][1]
    x = (lo - 0.02 * (hi - lo) + rng.rand(200000, 3) * 1.04 * (hi - lo)).astype(np.float32)
    fo = O.FieldOracle(H.oracle_spec(g), H.oracle_params(g))
    f_ref, c_ref = fo.get_coding(x)
    y_ref = O.mlp_forward(H._layers(g, 'param.linear_mat'), f_ref)
    f = _fused_fwd(m, G.t(x))
    assert H.rel_err(G.npy(f['feats']), f_ref) < 2e-5
    assert H.rel_err(G.npy(f['coeff']), c_ref) < 2e-5
    assert H.rel_err(G.npy(f['y']), y_ref) < 1e-4
