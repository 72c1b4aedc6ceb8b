import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import helpers as H, gpu_helpers as G
from ffb200 import native as nv
from ffb200.renderer import render_ray
g = H.golden('render_train')
for fused in (0, 1):
    nv.lib().ffb_set_fused_mlp(fused)
    cfg, m = G.build_model(g)
    S = int(g['N_samples'])
    m._jitter = lambda n, tr: G.t(g['jitter']) if tr else None
    out = render_ray(torch.from_numpy(g['rays']), m, chunk=4096, N_samples=S, white_bg=True, is_train=True, device='cuda')
    loss = torch.mean((out[0] - G.t(g['target'])) ** 2)
    params = list(m.named_parameters())
    grads = torch.autograd.grad(loss, [p for _, p in params], allow_unused=True)
    print('fused', fused, 'n_valid', int(m.last_stats['n_valid']), 'rgb', H.rel_err(G.npy(out[0]), g['rgb_map']))
    for (n, p), gr in zip(params, grads):
        print('   %-32s %.3e' % (n, H.rel_err(G.npy(gr), g['grad.' + n])))
