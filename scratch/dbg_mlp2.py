import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ffb200
from ffb200 import native as nv
lib = nv.lib()
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
for (K0, H, N) in [(18, 64, 32), (20, 128, 7)]:
    torch.manual_seed(0)
    n = 256 * 3
    x = torch.randn(n, K0, device='cuda'); W1 = torch.randn(H, K0, device='cuda') / K0 ** 0.5
    b1 = torch.randn(H, device='cuda') * 0.3; W2 = torch.randn(N, H, device='cuda') / H ** 0.5
    gy = torch.randn(n, N, device='cuda')
    gx = torch.zeros(n, K0, device='cuda'); gW1, gb1, gW2 = torch.zeros_like(W1), torch.zeros_like(b1), torch.zeros_like(W2)
    nv.check(lib.ffb_mlp2_bwd(P(x), P(gy), P(W1), P(b1), P(W2), P(gx), P(gW1), P(gb1), P(gW2), C.c_int64(n), None, K0, H, N, nv.stream()))
    xd, gd = x.double(), gy.double()
    h = torch.relu(xd @ W1.double().T + b1.double()); gh = (gd @ W2.double()) * (h > 0)
    for name, got, want in (('gx', gx, gh @ W1.double()), ('gW1', gW1, gh.T @ xd), ('gb1', gb1, gh.sum(0)), ('gW2', gW2, gd.T @ h)):
        e = (got.double() - want).abs()
        print((K0, H, N), name, 'relerr', float(e.max() / want.abs().max()), 'frac bad', float((e > 1e-3 * want.abs().max()).double().mean()))
    e = (gx.double() - gh @ W1.double()).abs()
    bad_rows = (e.max(1).values > 1e-3).nonzero().flatten()
    bad_cols = (e.max(0).values > 1e-3).nonzero().flatten()
    print('bad rows', bad_rows[:20].tolist(), len(bad_rows), 'bad cols', bad_cols.tolist())
    # try hypotheses: gx computed without relu mask?  with transposed W?
    alt = (gd @ W2.double()) @ W1.double()
    print('no-mask hypothesis err', float((gx.double() - alt).abs().max() / alt.abs().max()))
