"""Timeline of CTA 0 of the pipelined linear_mat forward kernel (ffb_mlp2p_trace)."""
import ctypes as C, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ffb200 import native as nv
lib = nv.lib()
n = 985030
torch.manual_seed(0)
x = torch.randn(n, 18, device='cuda'); W1 = torch.randn(64, 18, device='cuda') / 4; b1 = torch.randn(64, device='cuda') * .3
W2 = torch.randn(32, 64, device='cuda') / 8; y = torch.empty(n, 32, device='cuda'); bits = torch.empty(n, 4, device='cuda', dtype=torch.int16)
P = lambda t: C.c_void_p(t.data_ptr())
run = lambda: nv.check(lib.ffb_mlp2_fwd(P(x), P(W1), P(b1), P(W2), P(y), P(bits), C.c_int64(n), None, 18, 64, 32, nv.stream()))
for _ in range(3): run()
buf = torch.zeros(60000, dtype=torch.int64, device='cuda')
nv.check(lib.ffb_mlp2p_trace(C.c_void_p(buf.data_ptr())))
run(); torch.cuda.synchronize()
nv.check(lib.ffb_mlp2p_trace(None))
ev = buf.view(-1, 3).cpu().numpy()
ev = ev[ev[:, 2] > 0]
t0 = ev[:, 2].min()
names = {40: 'I slot_full', 41: 'I acc_free', 42: 'I issued', 50: 'E epi2 start', 51: 'E L2 complete', 52: 'E epi2 done', 53: 'E L1 complete', 54: 'E epi1 done', 55: 'E synced'}
rows = sorted((int(e[2] - t0), int(e[1]), int(e[0])) for e in ev)
for t, tile, e in rows:
    if 20 <= tile <= 23 and e >= 40:
        nm = names.get(e, ('P%d loaded' % (e - 10)) if e < 20 else ('P%d slot_free' % (e - 20)) if e < 30 else ('P%d delivered' % (e - 30)))
        print(f'{t:8d} ns  tile {tile:3d}  {nm}')
print('events', len(rows), 'span us', (ev[:, 2].max() - t0) / 1e3)
