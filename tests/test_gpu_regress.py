"""GPU: the regression drivers (ffb200.train.regression / RegressStep — scripts/2D_regression.ipynb, sdf_regression.ipynb,
2D_set_regression.py of the reference) against the reference's step restated with torch CPU operators
(oracle/torch_port.RegressPort, pinned to the golden field vectors): identical initial weights and identical index
draws (torch.randint on the CPU generator) on both sides; the loss curve and the end-of-run PSNR must agree."""
import numpy as np
import pytest
import torch

from tests import synth_scene as SS

pytestmark = pytest.mark.gpu

SMALL_IMG = ['model.basis_dims=[8,8,8,4,4,4]', 'model.basis_resos=[8,13,18,22,27,32]', 'model.total_params=40000']


def _psnr(mse):
    return -10.0 * np.log10(max(float(mse), 1e-12))


def _run_pair(cfgname, overrides, aabb, coords, targets, steps, batch, scale_loss, coef_mode='bilinear', basis_mode='bilinear',
              use_graph=True):
    import ffb200
    from ffb200.models.FactorFields import FactorFields
    from ffb200.train import evaluate_field, regression
    from oracle.torch_port import RegressPort
    cfg = ffb200.load_cfg(cfgname, overrides + [f'training.n_iters={steps}', f'training.batch_size={batch}'])
    cfg.dataset.aabb = aabb
    torch.manual_seed(3)
    m = FactorFields(cfg, 'cuda:0')
    state = {k: v.detach().cpu().contiguous().numpy() for k, v in m.state_dict().items()}
    ct, tt = torch.from_numpy(coords), torch.from_numpy(targets)
    N = ct.shape[0]
    # ours
    torch.manual_seed(17)
    res = regression(cfg, m, ct, tt, n_iters=steps, batch_size=batch, scale_loss=scale_loss, use_graph=use_graph)
    pred = evaluate_field(m, ct).cpu()
    ours_final = _psnr(torch.mean((pred - tt) ** 2))
    # reference port, same index stream
    torch.manual_seed(17)
    decay = 0.1 ** (1.0 / steps) if scale_loss else 1.0
    rp = RegressPort(state, m.aabb.cpu().numpy(), m.freq_bands.cpu().numpy(), m.in_dim, coef_mode, basis_mode,
                     lr_small=cfg.training.lr_small, lr_large=cfg.training.lr_large, loss_scale_decay=decay)
    ref_loss = []
    for it in range(steps):
        idx = torch.randint(0, N, (batch,))
        ref_loss.append(rp.train_step(ct[idx], tt[idx]))
    ref_pred = torch.cat([rp.predict(c) for c in torch.split(ct, 16384)])
    ref_final = _psnr(torch.mean((ref_pred - tt) ** 2))
    return res, ours_final, np.array(ref_loss), ref_final


def _check(tag, res, ours_final, ref_loss, ref_final, min_gain_db):
    ours = np.array(res['loss'])
    print(f'{tag}: first loss ours {ours[0]:.6e} ref {ref_loss[0]:.6e}; last-20 PSNR ours {_psnr(ours[-20:].mean()):.3f} ref '
          f'{_psnr(ref_loss[-20:].mean()):.3f}; full-set PSNR ours {ours_final:.3f} ref {ref_final:.3f} dB')
    assert abs(ours[0] - ref_loss[0]) <= 2e-5 * abs(ref_loss[0])            # same weights, same batch
    assert _psnr(ref_loss[-20:].mean()) - _psnr(ref_loss[0]) > min_gain_db, 'nothing was learnt: the comparison would be vacuous'
    assert abs(_psnr(ours[-20:].mean()) - _psnr(ref_loss[-20:].mean())) < 0.1
    assert abs(ours_final - ref_final) < 0.1                                  # north_star: end-of-run PSNR within 0.1 dB


@pytest.mark.parametrize('use_graph', [True, False])
def test_image_regression_parity(use_graph):
    """image.yaml at reduced size (nearest-mode coefficient and basis grids, 36 -> 64 -> 3 MLP), 128 x 128 synthetic image."""
    coords, img = SS.synth_image(128, 128, 3, seed=0)
    out = _run_pair('image.yaml', SMALL_IMG, [[0., 0.], [128., 128.]], coords, img, steps=300, batch=4096, scale_loss=True,
                    coef_mode='nearest', basis_mode='nearest', use_graph=use_graph)
    _check('image', *out, min_gain_db=8.0)


def test_sdf_regression_parity():
    """sdf.yaml at reduced size (3-D trilinear grids, bias-free Linear(18, 1)), analytic SDF samples."""
    coords, sdf = SS.synth_sdf(60000, 96.0, seed=1)
    out = _run_pair('sdf.yaml', ['model.total_params=40000'], [[0., 0., 0.], [96., 96., 96.]], coords, sdf, steps=300, batch=4096,
                    scale_loss=True)
    _check('sdf', *out, min_gain_db=6.0)


def test_image_set_regression_parity():
    """image_set.yaml at reduced size (one coefficient slab per image, shared bilinear bases), dropout off for parity."""
    coords, imgs = SS.synth_image_set(6, 32, 32, seed=2)
    out = _run_pair('image_set.yaml', SMALL_IMG + ['model.with_dropout=false'], [[0, 0, 0], [32, 32, 6]], coords, imgs, steps=300,
                    batch=2048, scale_loss=False)
    _check('image_set', *out, min_gain_db=5.0)


def test_image_set_regression_with_dropout_runs():
    """with_dropout=true (the shipped image_set.yaml): F.dropout(p=0.1) on the MLP input when is_train (FactorFields.py:150-151)
    draws from the device RNG inside the captured step: replays must draw fresh masks, and the run must still learn."""
    import ffb200
    from ffb200.models.FactorFields import FactorFields
    from ffb200.train import RegressStep, regression
    coords, imgs = SS.synth_image_set(6, 32, 32, seed=2)
    cfg = ffb200.load_cfg('image_set.yaml', SMALL_IMG + ['training.n_iters=200', 'training.batch_size=2048'])
    cfg.dataset.aabb = [[0, 0, 0], [32, 32, 6]]
    torch.manual_seed(3)
    m = FactorFields(cfg, 'cuda:0')
    with torch.no_grad():
        for p in m.coeffs:
            p.add_(0.5)             # coef_init 0.001 makes the features ~0: lift them so a dropped input is visible in the loss
    # lr 0: replays leave the weights alone, so the loss of the same batch changes only through the dropout mask
    rs = RegressStep(m, m.get_optparam_groups(0.0, 0.0), batch=2048, x_dim=3, out_dim=3, is_train=True)
    x, t = torch.from_numpy(coords[:2048]).cuda(), torch.from_numpy(imgs[:2048]).cuda()
    l = [float(rs.step(x, t).item()) for _ in range(3)]
    assert len(set(l)) == 3, l
    torch.manual_seed(3)
    m = FactorFields(cfg, 'cuda:0')
    res = regression(cfg, m, torch.from_numpy(coords), torch.from_numpy(imgs), n_iters=200, batch_size=2048)
    loss = np.array(res['loss'])
    assert _psnr(loss[-20:].mean()) - _psnr(loss[0]) > 4.0
