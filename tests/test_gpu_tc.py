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


@pytest.mark.parametrize('K0,N,cap,use_ndev,use_bits', [(18, 32, 300077, False, True), (18, 32, 300077, True, True), (18, 32, 130001, False, False),
                                                        (20, 7, 70001, True, True), (31, 32, 1000, False, True), (5, 1, 129, False, True)])
def test_pipelined_mlp2_many_tiles(K0, N, cap, use_ndev, use_bits):
    """The pipelined warp-specialised linear_mat kernels (mlp_pipe.cu) with many tiles per CTA (ring wrap-around, both
    accumulator buffers, weight gradients resident in TMEM across tiles) against float64; and against the one-tile-at-a-time
    kernels of mlp_fused.cu on the same inputs."""
    from ffb200 import native as nv
    lib = nv.lib()
    H = 64
    assert lib.ffb_mlp2_pipelined_eligible(K0, H, N) == 1
    torch.manual_seed(K0 * 7 + N)
    n = cap - 1500 if use_ndev else cap
    if n <= 0:
        n = cap
    x = torch.randn(cap, K0, device='cuda')
    W1 = torch.randn(H, K0, device='cuda') / K0 ** 0.5
    b1 = torch.randn(H, device='cuda') * 0.3
    W2 = torch.randn(N, H, device='cuda') / H ** 0.5
    gy = torch.randn(cap, N, device='cuda')
    for _ in range(50):
        near = ((x.double() @ W1.double().T + b1.double()).abs() < 1e-3).any(1)
        if not bool(near.any()):
            break
        x[near] = torch.randn(int(near.sum()), K0, device='cuda')
    n_dev = torch.tensor([n], device='cuda', dtype=torch.int32) if use_ndev and n != cap else None
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    s = nv.stream()
    res = {}
    for pipelined in (1, 0):
        lib.ffb_set_mlp_pipelined(pipelined)
        try:
            if not pipelined and not lib.ffb_mlp2_eligible(K0, H, N):
                continue
            y = torch.full((cap, N), 7.0, device='cuda')
            bits = torch.zeros(cap, H // 16, device='cuda', dtype=torch.int16)
            nv.check(lib.ffb_mlp2_fwd(P(x), P(W1), P(b1), P(W2), P(y), P(bits), C.c_int64(cap), P(n_dev), K0, H, N, s))
            gx = torch.full((cap, K0), 7.0, device='cuda')
            gW1, gb1, gW2 = torch.zeros_like(W1), torch.zeros_like(b1), torch.zeros_like(W2)
            nv.check(lib.ffb_mlp2_bwd(P(x), P(gy), P(W1), P(b1), P(W2), P(bits) if use_bits else None, P(gx), P(gW1), P(gb1), P(gW2),
                                      C.c_int64(cap), P(n_dev), K0, H, N, s))
            torch.cuda.synchronize()
            res[pipelined] = (y, bits, gx, gW1, gb1, gW2)
        finally:
            lib.ffb_set_mlp_pipelined(1)
    y, bits, gx, gW1, gb1, gW2 = res[1]
    xd, gd = x[:n].double(), gy[:n].double()
    h = torch.relu(xd @ W1.double().T + b1.double())
    ref = h @ W2.double().T
    assert float((y[:n].double() - ref).abs().max() / ref.abs().max()) < 3e-6
    assert bool((y[n:] == 7.0).all()), 'rows beyond the device-side count were written'
    got_bits = ((bits[:n].int() & 0xffff)[:, :, None] >> torch.arange(16, device='cuda')) & 1
    assert bool((got_bits.reshape(n, H).bool() == (h > 0)).all()), 'ReLU decision bits'
    gh = (gd @ W2.double()) * (h > 0)
    for name, got, want in (('gx', gx[:n], gh @ W1.double()), ('gW1', gW1, gh.T @ xd), ('gb1', gb1, gh.sum(0)), ('gW2', gW2, gd.T @ h)):
        err = float((got.double() - want).abs().max() / want.abs().max())
        assert err < 3e-5, (name, err)
    assert bool((gx[n:] == 7.0).all())
    if 0 in res:
        for a, b in zip(res[1], res[0]):
            if a.dtype == torch.int16:
                assert torch.equal(a[:n], b[:n])
            else:
                assert float((a.double() - b.double())[:n if a.shape[0] == cap else None].abs().max() / b.double().abs().max()) < 3e-5


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


def _rgb_reference(feat, viewdirs, W1, b1, W2, b2, W3, view_pe, fea_pe):
    """MLPRender_Fea.forward (FactorFields.py:188-203) in float64."""
    def pe(p, freqs):
        fb = 2.0 ** torch.arange(freqs, dtype=torch.float64, device=p.device)
        pts = (p[..., None] * fb).reshape(p.shape[0], -1)
        return torch.cat([torch.sin(pts), torch.cos(pts)], -1)
    f, v = feat.double(), viewdirs.double()
    x = torch.cat([f, v, pe(f, fea_pe), pe(v, view_pe)], -1)
    h1 = torch.relu(x @ W1.double().T + b1.double())
    h2 = torch.relu(h1 @ W2.double().T + b2.double())
    return x, h1, h2, torch.sigmoid(h2 @ W3.double().T)


@pytest.mark.parametrize('n,use_idx', [(128 * 7 + 5, True), (40000, True), (3000, False)])
def test_fused_appearance_mlp_forward(n, use_idx):
    """mlp_rgb.cu: gather + input assembly + 3 layers + sigmoid in one tcgen05 kernel vs a float64 restatement of
    MLPRender_Fea; the fp32 activation copies and the ReLU decision bits it hands to the backward pass as well."""
    from ffb200 import native as nv
    lib = nv.lib()
    Cf, Hd, vpe, fpe = 31, 128, 6, 2
    K0 = 3 + Cf + 6 * vpe + 2 * fpe * Cf
    wsb = int(lib.ffb_rgbmlp_workspace_bytes(Cf, Hd, vpe, fpe))
    assert wsb > 0
    torch.manual_seed(n)
    Nv, R = (3 * n, 500) if use_idx else (n, n)
    feat = torch.randn(Nv, Cf + 1, device='cuda')
    rays = torch.randn(R, 6, device='cuda')
    rays[:, 3:] = torch.nn.functional.normalize(rays[:, 3:], dim=-1)
    ray_id = torch.randint(0, R, (Nv,), device='cuda', dtype=torch.int32) if use_idx else None
    app_idx = torch.sort(torch.randperm(Nv, device='cuda')[:n])[0].to(torch.int32) if use_idx else None
    W1 = torch.randn(Hd, K0, device='cuda') / K0 ** 0.5
    b1 = torch.randn(Hd, device='cuda') * 0.1
    W2 = torch.randn(Hd, Hd, device='cuda') / Hd ** 0.5
    b2 = torch.randn(Hd, device='cuda') * 0.1
    W3 = torch.randn(3, Hd, device='cuda') / Hd ** 0.5
    ws = torch.empty(wsb, device='cuda', dtype=torch.uint8)
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    rgb = torch.full((n, 3), -1.0, device='cuda')
    bits = torch.zeros(n, 16, device='cuda', dtype=torch.int16)
    x_o, h1_o, h2_o = torch.zeros(n, K0, device='cuda'), torch.zeros(n, Hd, device='cuda'), torch.zeros(n, Hd, device='cuda')
    s = nv.stream()
    _call(lib, 'ffb_rgbmlp_pack', P(W1), P(b1), P(W2), P(W3), P(ws), Cf, vpe, fpe, s)
    _call(lib, 'ffb_rgbmlp_fwd', P(feat), Cf + 1, P(rays), P(ray_id), P(app_idx), P(ws), P(b2), P(rgb), P(bits), P(x_o), P(h1_o), P(h2_o),
          None, None, None, C.c_int64(n), None, Cf, vpe, fpe, s)
    torch.cuda.synchronize()
    sel = app_idx.long() if use_idx else torch.arange(n, device='cuda')
    vd = rays[ray_id.long()[sel], 3:] if use_idx else rays[:, 3:]
    x, h1, h2, ref = _rgb_reference(feat[sel, 1:], vd, W1, b1, W2, b2, W3, vpe, fpe)
    rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max())
    assert rel(x_o, x) < 2e-6, rel(x_o, x)
    assert rel(h1_o, h1) < 3e-6, rel(h1_o, h1)
    assert rel(h2_o, h2) < 3e-6, rel(h2_o, h2)
    assert rel(rgb, ref) < 3e-6, rel(rgb, ref)
    # decision bits == sign of the activations the kernel itself produced
    shifts = torch.arange(16, device='cuda', dtype=torch.int32)
    unpack = lambda w: ((w.to(torch.int32)[..., None] >> shifts) & 1).reshape(n, -1).bool()
    assert torch.equal(unpack(bits[:, :8]), h1_o > 0) and torch.equal(unpack(bits[:, 8:]), h2_o > 0)


@pytest.mark.parametrize('n', [128 * 3 + 17, 30000])
def test_fused_appearance_mlp_backward(n):
    """mlp_rgb.cu backward (operand streams from the forward kernel, weight gradients accumulated in TMEM) vs float64
    autograd of MLPRender_Fea: input gradient and all five parameter gradients."""
    from ffb200 import native as nv
    lib = nv.lib()
    Cf, Hd, vpe, fpe = 31, 128, 6, 2
    K0 = 3 + Cf + 6 * vpe + 2 * fpe * Cf
    torch.manual_seed(n + 1)
    feat = torch.randn(n, Cf + 1, device='cuda')
    rays = torch.randn(n, 6, device='cuda')
    W1 = torch.randn(Hd, K0, device='cuda') / K0 ** 0.5
    b1 = torch.randn(Hd, device='cuda') * 0.1
    W2 = torch.randn(Hd, Hd, device='cuda') / Hd ** 0.5
    b2 = torch.randn(Hd, device='cuda') * 0.1
    W3 = torch.randn(3, Hd, device='cuda') / Hd ** 0.5
    g_rgb = torch.randn(n, 3, device='cuda')
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    ws = torch.empty(int(lib.ffb_rgbmlp_workspace_bytes(Cf, Hd, vpe, fpe)), device='cuda', dtype=torch.uint8)
    nbx, nbh = (int(lib.ffb_rgbmlp_stream_bytes(Cf, vpe, fpe, C.c_int64(n), w)) for w in (0, 1))
    sx, sh1, sh2 = (torch.empty(b, device='cuda', dtype=torch.uint8) for b in (nbx, nbh, nbh))
    rgb = torch.empty(n, 3, device='cuda')
    bits = torch.zeros(n, 16, device='cuda', dtype=torch.int16)
    s = nv.stream()
    _call(lib, 'ffb_rgbmlp_pack', P(W1), P(b1), P(W2), P(W3), P(ws), Cf, vpe, fpe, s)
    _call(lib, 'ffb_rgbmlp_fwd', P(feat), Cf + 1, P(rays), None, None, P(ws), P(b2), P(rgb), P(bits), None, None, None, P(sx), P(sh1), P(sh2),
          C.c_int64(n), None, Cf, vpe, fpe, s)
    ld_gx = (K0 + 3) // 4 * 4
    g_x = torch.zeros(n, ld_gx, device='cuda')
    gW1, gb1, gW2, gb2, gW3 = (torch.zeros_like(t) for t in (W1, b1, W2, b2, W3))
    _call(lib, 'ffb_rgbmlp_bwd', P(g_rgb), P(rgb), P(bits), P(sx), P(sh1), P(sh2), P(ws), P(W3), P(g_x), ld_gx, P(gW1), P(gb1), P(gW2), P(gb2),
          P(gW3), C.c_int64(n), None, Cf, vpe, fpe, s)
    g_x = g_x[:, :K0]
    torch.cuda.synchronize()
    # float64 reference of the backward pass, with the ReLU decisions the forward kernel recorded (a hidden unit within
    # rounding distance of zero may be decided differently in float64; its sample's g_x row then differs by a whole term)
    x, h1, h2, y = _rgb_reference(feat[:, 1:], rays[:, 3:], W1, b1, W2, b2, W3, vpe, fpe)
    shifts = torch.arange(16, device='cuda', dtype=torch.int32)
    unpack = lambda w: ((w.to(torch.int32)[..., None] >> shifts) & 1).reshape(n, -1).double()
    m1, m2 = unpack(bits[:, :8]), unpack(bits[:, 8:])
    assert float((m1 != (h1 > 0)).double().mean()) < 1e-4 and float((m2 != (h2 > 0)).double().mean()) < 1e-4
    g3 = g_rgb.double() * y * (1 - y)
    G2 = (g3 @ W3.double()) * m2
    G1 = (G2 @ W2.double()) * m1
    ref = [G1 @ W1.double(), G1.T @ x, G1.sum(0), G2.T @ h1, G2.sum(0), g3.T @ h2]
    rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max())
    for name, got, want in zip(['g_x', 'gW1', 'gb1', 'gW2', 'gb2', 'gW3'], [g_x, gW1, gb1, gW2, gb2, gW3], ref):
        assert rel(got, want) < 5e-5, (name, rel(got, want))


@pytest.mark.parametrize('K0,N,cap,use_ndev', [(18, 32, 200077, False), (18, 32, 200077, True), (20, 8, 5000, True)])
def test_pipelined_mlp2_backward_sparse_upstream_gradient(K0, N, cap, use_ndev):
    """ffb_mlp2p_bwd_sparse (the render path's hand-off: a density gradient for every row, feature gradients only for the rows with
    a slot in a compact table) against ffb_mlp2_bwd on the equivalent dense gradient — same operands, so the results agree to the
    rounding of the cross-CTA accumulation; rows without a slot must not read the compact table."""
    from ffb200 import native as nv
    lib = nv.lib()
    H = 64
    assert lib.ffb_mlp2_pipelined_eligible(K0, H, N) == 1
    torch.manual_seed(K0 + N + cap)
    n = cap - 777 if use_ndev else cap
    x = torch.randn(cap, K0, device='cuda')
    W1 = torch.randn(H, K0, device='cuda') / K0 ** 0.5
    b1 = torch.randn(H, device='cuda') * 0.3
    W2 = torch.randn(N, H, device='cuda') / H ** 0.5
    n_dev = torch.tensor([n], device='cuda', dtype=torch.int32) if use_ndev else None
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    s = nv.stream()
    y = torch.empty(cap, N, device='cuda')
    bits = torch.zeros(cap, H // 16, device='cuda', dtype=torch.int16)
    nv.check(lib.ffb_mlp2_fwd(P(x), P(W1), P(b1), P(W2), P(y), P(bits), C.c_int64(cap), P(n_dev), K0, H, N, s))
    # ~15 % of the rows carry feature gradients, in a shuffled compact table padded with NaN rows that nothing may read
    shaded = torch.rand(cap, device='cuda') < 0.15
    idx = torch.nonzero(shaded).flatten()
    n_rows = idx.numel()
    slot = torch.full((cap,), -1, device='cuda', dtype=torch.int32)
    slot[idx] = torch.arange(n_rows, device='cuda', dtype=torch.int32)
    rows = torch.full((n_rows + 64, N), float('nan'), device='cuda')
    rows[:n_rows] = torch.randn(n_rows, N, device='cuda')
    rows[:n_rows, 0] = float('nan')                       # column 0 of the compact rows is ignored
    g0 = torch.randn(cap, device='cuda')
    gy = torch.zeros(cap, N, device='cuda')
    gy[idx] = rows[:n_rows]
    gy[:, 0] = g0
    out = []
    for sparse in (False, True):
        gx = torch.full((cap, K0), 7.0, device='cuda')
        gW1, gb1, gW2 = torch.zeros_like(W1), torch.zeros_like(b1), torch.zeros_like(W2)
        if sparse:
            nv.check(lib.ffb_mlp2p_bwd_sparse(P(x), P(g0), P(slot), P(rows), P(W1), P(b1), P(W2), P(bits), P(gx), P(gW1), P(gb1), P(gW2),
                                              C.c_int64(cap), P(n_dev), K0, H, N, s))
        else:
            nv.check(lib.ffb_mlp2_bwd(P(x), P(gy), P(W1), P(b1), P(W2), P(bits), P(gx), P(gW1), P(gb1), P(gW2), C.c_int64(cap), P(n_dev),
                                      K0, H, N, s))
        out.append((gx, gW1, gb1, gW2))
    assert torch.equal(out[0][0][:n], out[1][0][:n])                       # input gradient: row-local arithmetic, bit-identical
    assert bool((out[1][0][n:] == 7.0).all())
    for name, a, b in zip(('gW1', 'gb1', 'gW2'), out[0][1:], out[1][1:]):
        assert bool(torch.isfinite(b).all()), name
        assert float((a - b).abs().max()) <= 2e-6 * float(a.abs().max()), name
