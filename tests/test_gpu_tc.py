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
