"""Drive the pipelined linear_mat kernels once each at the bench size (for ncu)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ffb200 import native as nv
lib = nv.lib()
n = 985030
torch.manual_seed(0)
x = torch.randn(n, 18, device='cuda'); W1 = torch.randn(64, 18, device='cuda') / 4; b1 = torch.randn(64, device='cuda') * .3
W2 = torch.randn(32, 64, device='cuda') / 8; y = torch.empty(n, 32, device='cuda'); bits = torch.empty(n, 4, device='cuda', dtype=torch.int16)
gy = torch.randn(n, 32, device='cuda'); gx = torch.empty(n, 18, device='cuda')
gW1, gb1, gW2 = torch.zeros_like(W1), torch.zeros_like(b1), torch.zeros_like(W2)
P = lambda t: C.c_void_p(t.data_ptr())
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    nv.check(lib.ffb_mlp2_fwd(P(x), P(W1), P(b1), P(W2), P(y), P(bits), C.c_int64(n), None, 18, 64, 32, nv.stream()))
    nv.check(lib.ffb_mlp2_bwd(P(x), P(gy), P(W1), P(b1), P(W2), P(bits), P(gx), P(gW1), P(gb1), P(gW2), C.c_int64(n), None, 18, 64, 32, nv.stream()))
torch.cuda.synchronize()
print('ok')
