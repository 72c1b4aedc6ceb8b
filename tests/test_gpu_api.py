"""GPU: the free functions and small methods of the module surface against vectors recorded from the reference
(tests/golden/api.npz): positional_encoding, raw2alpha, basis2density, normalize_basis, get_optparam_groups, n_parameters, and
the checkpoint layout of save() / load()."""
import json

import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_module_surface_golden(tmp_path):
    from ffb200.models.FactorFields import AlphaGridMask, positional_encoding, raw2alpha
    from tests import gpu_helpers as G
    g = H.golden('api')
    assert H.rel_err(G.npy(positional_encoding(G.t(g['pe_x']), 4)), g['pe_y']) < 2e-6
    a, w, bg = raw2alpha(G.t(g['r2a_sigma']), G.t(g['r2a_dist']))
    assert H.rel_err(G.npy(a), g['r2a_alpha']) < 1e-5 and H.rel_err(G.npy(w), g['r2a_weight']) < 1e-5
    assert H.rel_err(G.npy(bg), g['r2a_bg']) < 1e-5
    cfg, m = G.build_model(g)
    f = G.t(g['b2d_f'])
    assert H.rel_err(G.npy(m.basis2density(f)), g['b2d_softplus']) < 1e-6
    cfg.renderer.fea2denseAct = 'relu'
    assert H.rel_err(G.npy(m.basis2density(f)), g['b2d_relu']) < 1e-6
    cfg.renderer.fea2denseAct = 'softplus'
    groups = m.get_optparam_groups(lr_small=0.001, lr_large=0.02)
    assert [[gr['lr'], [list(p.shape) for p in gr['params']]] for gr in groups] == json.loads(str(g['groups']))
    assert m.n_parameters() == int(g['fact.n_parameters'])
    # checkpoint: same keys / tensor names / shapes / packed mask bytes as the reference's save() (FactorFields.py:662-679)
    m.alphaMask = AlphaGridMask('cuda', m.aabb, G.t(g['ck_volume']))
    path = str(tmp_path / 'ck.th')
    m.save(path)
    ck = torch.load(path, weights_only=False)
    assert sorted(ck.keys()) == json.loads(str(g['ck_keys']))
    assert [[k, list(v.shape)] for k, v in ck['state_dict'].items()] == json.loads(str(g['ck_state_keys']))
    assert all(v.is_contiguous() for v in ck['state_dict'].values())          # the reference's (channel-first) storage order on disk
    for k, v in ck['state_dict'].items():
        assert np.array_equal(v.cpu().numpy(), g['param.' + k]), k
    assert np.array_equal(ck['alphaMask.mask'], g['ck_mask']) and list(ck['alphaMask.shape']) == g['ck_mask_shape'].tolist()
    assert np.allclose(ck['alphaMask.aabb'].cpu().numpy(), g['ck_mask_aabb'])
    # load() into a fresh model restores the parameters and the mask (FactorFields.py:681-691)
    cfg2, m2 = G.build_model(g)
    with torch.no_grad():
        for p in m2.parameters():
            p.zero_()
    m2.load(ck)
    for (n, p), q in zip(m.named_parameters(), m2.parameters()):
        assert torch.equal(p, q), n
    assert torch.equal(m2.alphaMask.alpha_volume, m.alphaMask.alpha_volume)
    # normalize_basis (:518-521)
    with torch.no_grad():
        m.normalize_basis()
    assert H.rel_err(G.npy(m.basises[0]), g['normalized_basis0']) < 1e-5 and H.rel_err(G.npy(m.basises[5]), g['normalized_basis5']) < 1e-5
    x = G.t(g['pe_x'][:, :3]) * 0.5
    feats, _ = m.get_coding(x)              # the re-normalised (re-allocated) bases are what the kernels read
    assert bool(torch.isfinite(feats).all())
