"""Times the fused appearance-MLP kernels alone (CUDA events), n = 147k shaded samples as in the nerf.yaml bench step."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ffb200 import native as nv
lib = nv.lib()
Cf, Hd, vpe, fpe = 31, 128, 6, 2
K0 = 194
n = int(sys.argv[1]) if len(sys.argv) > 1 else 147000
Nv, R = 990000, 4096
torch.manual_seed(0)
feat = torch.randn(Nv, Cf + 1, device='cuda')
rays = torch.randn(R, 6, device='cuda')
ray_id = torch.sort(torch.randint(0, R, (Nv,), device='cuda'))[0].to(torch.int32)
app_idx = torch.sort(torch.randperm(Nv, device='cuda')[:n])[0].to(torch.int32)
W1 = torch.randn(Hd, K0, device='cuda') / K0 ** 0.5; b1 = torch.randn(Hd, device='cuda') * 0.1
W2 = torch.randn(Hd, Hd, device='cuda') / Hd ** 0.5; b2 = torch.randn(Hd, device='cuda') * 0.1
W3 = torch.randn(3, Hd, device='cuda') / Hd ** 0.5
ws = torch.empty(int(lib.ffb_rgbmlp_workspace_bytes(Cf, Hd, vpe, fpe)), device='cuda', dtype=torch.uint8)
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
rgb = torch.empty(n, 3, device='cuda'); bits = torch.empty(n, 16, device='cuda', dtype=torch.int16)
x_o, h1_o, h2_o = torch.empty(n, K0, device='cuda'), torch.empty(n, Hd, device='cuda'), torch.empty(n, Hd, device='cuda')
s = nv.stream()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print('pack        %.1f us' % timeit(lambda: nv.check(lib.ffb_rgbmlp_pack(P(W1), P(b1), P(W2), P(W3), P(ws), Cf, vpe, fpe, s))))
nbx, nbh = (int(lib.ffb_rgbmlp_stream_bytes(Cf, vpe, fpe, C.c_int64(n), w)) for w in (0, 1))
sx, sh1, sh2 = (torch.empty(b, device='cuda', dtype=torch.uint8) for b in (nbx, nbh, nbh))
for tag, outs in [('fwd (rgb only)', (None,) * 7), ('fwd + bits', (bits,) + (None,) * 6), ('fwd + bits + fp32 x/h1/h2', (bits, x_o, h1_o, h2_o, None, None, None)),
                  ('fwd + bits + bf16 streams', (bits, None, None, None, sx, sh1, sh2))]:
    f = lambda: nv.check(lib.ffb_rgbmlp_fwd(P(feat), Cf + 1, P(rays), P(ray_id), P(app_idx), P(ws), P(b2), P(rgb), *[P(o) for o in outs],
                                            C.c_int64(n), None, Cf, vpe, fpe, s))
    print('%-28s %.1f us  (%.1f us per 128-row tile per SM)' % (tag, timeit(f), timeit(f) / max(1, (n + 127) // 128 / 148)))
g_rgb = torch.randn(n, 3, device='cuda')
g_x = torch.empty(n, 196, device='cuda')
gW1, gb1, gW2, gb2, gW3 = (torch.zeros_like(t) for t in (W1, b1, W2, b2, W3))
fb = lambda: nv.check(lib.ffb_rgbmlp_bwd(P(g_rgb), P(rgb), P(bits), P(sx), P(sh1), P(sh2), P(ws), P(W3), P(g_x), 196, P(gW1), P(gb1), P(gW2), P(gb2),
                                         P(gW3), C.c_int64(n), None, Cf, vpe, fpe, s))
print('%-28s %.1f us' % ('bwd', timeit(fb)))
