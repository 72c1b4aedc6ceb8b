"""Is the appearance-MLP weight-gradient error of render_ndc_eval_alpha due to ReLU decisions flipping under the 2-part bf16 split?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import helpers as H, gpu_helpers as G
import ffb200
from ffb200 import ops
from ffb200.models.FactorFields import AlphaGridMask
from ffb200.renderer import render_ray
for name in ['ndc_eval_alpha', 'train', 'eval_alpha', 'unbound_train']:
    g = H.golden('render_' + name)
    for terms in (2, 3):
        ops.APPEARANCE_TERMS = terms
        cfg, m = G.build_model(g)
        if 'alpha_volume' in g:
            m.alphaMask = AlphaGridMask('cuda', G.t(g['alpha_aabb']), G.t(g['alpha_volume']))
        is_train = bool(g['is_train']); mode = str(g['mode']) if 'mode' in g else 'bounded'
        m._jitter = lambda n, tr: G.t(g['jitter']) if tr else None
        m._z_uniform = lambda n, tr: torch.from_numpy(g['jitter']) if tr else None
        out = render_ray(torch.from_numpy(g['rays']), m, chunk=4096, N_samples=int(g['N_samples']), ndc_ray=(mode == 'ndc'), white_bg=True, is_train=is_train, device='cuda')
        loss = torch.mean((out[0] - G.t(g['target'])) ** 2)
        params = list(m.named_parameters())
        grads = torch.autograd.grad(loss, [p for _, p in params], allow_unused=True)
        errs = {n: H.rel_err(G.npy(gr), g['grad.' + n]) for (n, p), gr in zip(params, grads) if gr is not None}
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
        print(name, 'terms', terms, 'rgb err', H.rel_err(G.npy(out[0]), g['rgb_map']), [(k, f'{v:.1e}') for k, v in worst], flush=True)
