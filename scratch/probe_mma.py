import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ffb200 import native as nv
out = torch.zeros(2, dtype=torch.int64, device='cuda')
for N in (16, 32, 64, 96, 128, 192, 256):
    for cnt in (64, 1024):
        nv.check(nv.lib().ffb_probe_mma(N, cnt, C.c_void_p(out.data_ptr()), nv.stream()))
        torch.cuda.synchronize()
        a, b = out.tolist()
        print(f'N={N:3d} count={cnt:5d}: issue {a / cnt:7.1f} cycles/MMA, issue+complete {b / cnt:7.1f} cycles/MMA  ({128 * N * 16 / (b / cnt):7.0f} MAC/cycle)')
