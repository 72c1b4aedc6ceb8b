"""tcgen05 tensor-core layer kernels (mlp_tc.cu) against the exact-fp32 SIMT kernels (linear.cu) and a float64
reference, through the C ABI."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(18, 64), (64, 32), (194, 128), (128, 128), (128, 3), (144, 64), (64, 1), (60, 256), (18, 1)]


def _call(lib, name, *args):
    from ffb200 import native as nv
    nv.check(getattr(lib, name)(*args))


@pytest.mark.parametrize('K,M', SHAPES)
@pytest.mark.parametrize('act', [0, 1, 2])
def test_tc_layers(K, M, act):
    from ffb200 import native as nv
    lib = nv.lib()
    if not lib.ffb_linear_tc_eligible(K, M):
        pytest.skip('shape not eligible')
    torch.manual_seed(K * 1000 + M + act)
    n = 5000 + 37
    x = torch.randn(n, K, device='cuda')
    W = torch.randn(M, K, device='cuda') / K ** 0.5
    b = torch.randn(M, device='cuda') * 0.1
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    s = nv.stream()
    # ---- forward
    y_tc = torch.empty(n, M, device='cuda')
    _call(lib, 'ffb_linear_tc_fwd', P(x), P(W), P(b), P(y_tc), C.c_int64(n), None, K, M, act, s)
    z = x.double() @ W.double().T + b.double()
    ref = z if act == 0 else torch.relu(z) if act == 1 else torch.sigmoid(z)
    err = float((y_tc.double() - ref).abs().max() / ref.abs().max())
    assert err < 3e-6, ('fwd', K, M, act, err)      # 3-term split: fp32-class accuracy
    # ---- input gradient: gx = (gy .* act'(y)) W
    gy = torch.randn(n, M, device='cuda')
    y = ref.float()
    mask = torch.ones_like(ref) if act == 0 else (ref > 0).double() if act == 1 else ref * (1 - ref)
    gm = gy.double() * mask
    if lib.ffb_linear_tc_eligible(M, K):
        gx_tc = torch.empty(n, K, device='cuda')
        _call(lib, 'ffb_linear_tc_bwd_input', P(gy), P(y), P(W), P(gx_tc), C.c_int64(n), None, K, M, act, s)
        gx_ref = gm @ W.double()
        err = float((gx_tc.double() - gx_ref).abs().max() / gx_ref.abs().max())
        assert err < 3e-5, ('dgrad', K, M, act, err)
    # ---- weight / bias gradient
    if lib.ffb_linear_tc_wgrad_eligible(K, M):
        gW = torch.zeros(M, K, device='cuda')
        gb = torch.zeros(M, device='cuda')
        _call(lib, 'ffb_linear_tc_bwd_weight', P(gy), P(y), act, P(x), P(gW), P(gb), C.c_int64(n), None, K, M, s)
        gW_ref = gm.T @ x.double()
        gb_ref = gm.sum(0)
        assert float((gW.double() - gW_ref).abs().max() / gW_ref.abs().max()) < 3e-5, ('wgrad', K, M, act)
        assert float((gb.double() - gb_ref).abs().max() / gb_ref.abs().max()) < 3e-5, ('bgrad', K, M, act)


def test_tc_large_and_device_count():
    """1 M rows of the linear_mat shapes + a device-side row count smaller than the launch bound."""
    from ffb200 import native as nv
    lib = nv.lib()
    n, K, M = 1 << 20, 18, 64
    x = torch.randn(n, K, device='cuda')
    W = torch.randn(M, K, device='cuda') * 0.2
    b = torch.randn(M, device='cuda') * 0.1
    y = torch.full((n, M), 7.0, device='cuda')
    n_dev = torch.tensor([n - 1000], device='cuda', dtype=torch.int32)
    P = lambda t: C.c_void_p(t.data_ptr())
    nv.check(lib.ffb_linear_tc_fwd(P(x), P(W), P(b), P(y), C.c_int64(n), P(n_dev), K, M, 1, nv.stream()))
    ref = torch.relu(x @ W.T + b)
    assert float((y[:n - 1000] - ref[:n - 1000]).abs().max()) < 1e-4
    assert bool((y[n - 1000:] == 7.0).all())


MLP2_SHAPES = [(18, 64, 32), (54, 64, 32), (20, 64, 7), (144, 64, 3), (33, 64, 16)]


@pytest.mark.parametrize('K0,H,N', MLP2_SHAPES)
@pytest.mark.parametrize('use_ndev', [False, True])
def test_fused_mlp2(K0, H, N, use_ndev):
    """Whole 2-layer MLPMixer in one tcgen05 kernel (mlp_fused.cu) against a float64 reference, through the C ABI:
    forward 3e-6 (3-part bf16 split), gradients 3e-5 (2-part split) relative to the largest reference magnitude."""
    from ffb200 import native as nv
    lib = nv.lib()
    if not lib.ffb_mlp2_eligible(K0, H, N):
        pytest.skip('shape not eligible')
    torch.manual_seed(K0 * 131 + H + N)
    cap = 5000 + 37
    n = cap - 1500 if use_ndev else cap
    x = torch.randn(cap, K0, device='cuda')
    W1 = torch.randn(H, K0, device='cuda') / K0 ** 0.5
    b1 = torch.randn(H, device='cuda') * 0.3
    W2 = torch.randn(N, H, device='cuda') / H ** 0.5
    gy = torch.randn(cap, N, device='cuda')
    # keep every hidden pre-activation away from 0, where the ReLU mask (hence the gradient) legitimately depends on
    # the last bits of the forward arithmetic
    for _ in range(50):
        near = ((x.double() @ W1.double().T + b1.double()).abs() < 1e-3).any(1)
        if not bool(near.any()):
            break
        x[near] = torch.randn(int(near.sum()), K0, device='cuda')
    n_dev = torch.tensor([n], device='cuda', dtype=torch.int32) if use_ndev else None
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    s = nv.stream()
    y = torch.full((cap, N), 7.0, device='cuda')
    bits = torch.zeros(cap, H // 16, device='cuda', dtype=torch.int16)
    nv.check(lib.ffb_mlp2_fwd(P(x), P(W1), P(b1), P(W2), P(y), P(bits), C.c_int64(cap), P(n_dev), K0, H, N, s))
    xd, gd = x[:n].double(), gy[:n].double()
    h = torch.relu(xd @ W1.double().T + b1.double())
    ref = h @ W2.double().T
    assert float((y[:n].double() - ref).abs().max() / ref.abs().max()) < 3e-6
    assert bool((y[n:] == 7.0).all()), 'rows beyond the device-side count were written'
    got_bits = ((bits[:n].int() & 0xffff)[:, :, None] >> torch.arange(16, device='cuda')) & 1
    assert bool((got_bits.reshape(n, H).bool() == (h > 0)).all()), 'ReLU decision bits'
    gx = torch.full((cap, K0), 7.0, device='cuda')
    gW1, gb1, gW2 = torch.zeros_like(W1), torch.zeros_like(b1), torch.zeros_like(W2)
    nv.check(lib.ffb_mlp2_bwd(P(x), P(gy), P(W1), P(b1), P(W2), P(bits) if K0 != 54 else None, P(gx), P(gW1), P(gb1), P(gW2), C.c_int64(cap),
                              P(n_dev), K0, H, N, s))
    gh = (gd @ W2.double()) * (h > 0)
    for name, got, want in (('gx', gx[:n], gh @ W1.double()), ('gW1', gW1, gh.T @ xd), ('gb1', gb1, gh.sum(0)), ('gW2', gW2, gd.T @ h)):
        err = float((got.double() - want).abs().max() / want.abs().max())
        assert err < 3e-5, (name, err)
    assert bool((gx[n:] == 7.0).all())


def test_fused_mlp2_matches_per_layer_path():
    """MLPMixer module: fused kernels vs the per-layer tcgen05 kernels on the nerf.yaml linear_mat shape (autograd)."""
    from ffb200 import native as nv
    from ffb200.models.FactorFields import MLPMixer
    torch.manual_seed(3)
    mm = MLPMixer(18, 32, num_layers=2, hidden_dim=64).cuda()
    x = torch.randn(20000, 18, device='cuda', requires_grad=True)
    G = torch.randn(20000, 32, device='cuda')
    outs = []
    for fused in (1, 0):
        nv.lib().ffb_set_fused_mlp(fused)
        y = mm(x)
        outs.append([y.detach()] + [g.detach() for g in torch.autograd.grad((y * G).sum(), [x] + list(mm.parameters()))])
    nv.lib().ffb_set_fused_mlp(1)
    for a, b in zip(*outs):
        assert float((a - b).abs().max() / b.abs().max()) < 3e-5
