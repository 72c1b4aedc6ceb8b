"""GPU parity of the alpha-mask maintenance path (FactorFields.py:693-841: compute_alpha, getDenseAlpha, updateAlphaMask,
filtering_rays, upsample_volume_grid, shrink) against vectors recorded from the unmodified reference
(tests/golden/maintenance.npz, see make_golden.maintenance_case)."""
import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_alpha_mask_maintenance_golden():
    from tests import gpu_helpers as G
    g = H.golden('maintenance')
    cfg, m = G.build_model(g)
    xyz = G.t(g['ca_xyz'])
    with torch.no_grad():
        # compute_alpha without a mask (:710-727)
        assert H.rel_err(G.npy(m.compute_alpha(xyz, length=0.2)), g['ca_alpha']) < 1e-4
        # dense lattice, no jitter (:730-755)
        gs = [20, 18, 22]
        alpha, dense_xyz = m.getDenseAlpha(gs, times=1)
        assert np.allclose(G.npy(dense_xyz).astype(np.float64).sum((0, 1, 2)), g['dense_xyz_sum'], rtol=1e-6)
        assert H.rel_err(G.npy(alpha), g['dense_alpha']) < 1e-4
        # mask update: max-pool, threshold, small-component filter, new box (:758-793)
        torch.manual_seed(777)       # the 16 jitter rounds draw torch.rand on the CPU generator (:748): same stream as the golden run
        new_aabb = m.updateAlphaMask(tuple(gs), is_update_alphaMask=True)
        vol = G.npy(m.alphaMask.alpha_volume[0, 0])
        assert (vol != g['mask_volume']).mean() < 2e-3               # a voxel whose pooled alpha sits on the threshold may flip
        assert np.allclose(G.npy(new_aabb), g['new_aabb'], atol=1e-6)
        assert np.allclose(G.npy(m.alphaMask.aabb), g['mask_aabb'])
        # compute_alpha through the mask (:712-716)
        assert H.rel_err(G.npy(m.compute_alpha(xyz, length=0.2)), g['ca_alpha_masked']) < 1e-4
        # ray filtering (:811-841): same rays kept, same order
        rays, rgbs = torch.from_numpy(g['f_rays'].copy()), torch.from_numpy(g['f_rgbs'].copy())
        r1, c1 = m.filtering_rays(rays.clone(), rgbs.clone(), N_samples=64, chunk=250)
        assert np.array_equal(r1.numpy(), g['f_kept_rays']) and np.array_equal(c1.numpy(), g['f_kept_rgbs'])
        r2, _ = m.filtering_rays(rays.clone(), rgbs.clone(), chunk=250, bbox_only=True)
        assert np.array_equal(r2.numpy(), g['f_kept_rays_bbox'])
        # render parameters after an upsample (:693-708)
        m.upsample_volume_grid([40, 36, 44])
        assert float(m.stepSize) == float(g['up_stepSize']) and m.nSamples == int(g['up_nSamples'])
        assert np.array_equal(G.npy(m.gridSize), g['up_gridSize'])
        # shrink re-initialises the factors at the new box (:795-809)
        m.shrink(G.t(g['new_aabb']))
        assert [int(v) for v in m.coeff_reso] == g['shrunk.coeff_reso'].tolist()
        assert list(m.basis_reso) == g['shrunk.basis_reso'].tolist()
        assert np.array_equal(G.npy(m.freq_bands), g['shrunk.freq_bands'])
        assert float(m.stepSize) == float(g['shrunk.stepSize']) and m.nSamples == int(g['shrunk.nSamples'])
        assert list(m.coeffs[0].shape) == g['shrunk.coeff_shape'].tolist()
        assert float(m.coeffs[0].flatten()[0]) == float(g['shrunk.coeff_const'])
        assert H.rel_err(G.npy(m.basises[0]), g['shrunk.basis0']) < 1e-6          # DCT re-initialisation
        assert np.allclose(np.array(cfg.dataset.aabb), g['shrunk.cfg_aabb'])
        assert m.n_parameters() == int(g['shrunk.n_parameters'])
        # and the re-initialised model still renders
        out = m(G.t(g['f_rays'][:64]), white_bg=True, is_train=False, N_samples=32)
        assert out[0].shape == (64, 3) and bool(torch.isfinite(out[0]).all())
