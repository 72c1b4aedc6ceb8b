# Baseline-measurement probe written at survey time (NOT product source, NOT a port).
# Usage (build session): mkdir -p baseline/_ref && cp -r /root/reference baseline/_ref/factor-fields
#                        gpurun --gpus 1 --timeout 900 -- python3 baseline/ref_probe.py image sdf nerf nerf_vm nerf_cp
# Times the UNMODIFIED reference (PyTorch eager) on the GPU box: CUDA path on one B200 and the
# CPU path on the host cores, at the BASELINE.json config shapes, on synthetic inputs.
# Also measures how far the reference's own CPU and CUDA paths disagree (calibrates parity bars).
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from refload import load_cfg
import torch, numpy as np
from models.FactorFields import FactorFields
import renderer as R
from utils import N_to_reso, cal_n_samples

out = {'env': {'nproc': os.cpu_count(), 'torch_threads': torch.get_num_threads(), 'torch': torch.__version__,
               'cuda': torch.cuda.is_available(),
               'gpu': torch.cuda.get_device_name(0) if torch.cuda.is_available() else None,
               'n_gpu': torch.cuda.device_count(),
               'allow_tf32_matmul': torch.backends.cuda.matmul.allow_tf32}}
try:
    out['env']['cpu_model'] = [l.split(':')[1].strip() for l in open('/proc/cpuinfo') if l.startswith('model name')][0]
except Exception:
    pass
HAS_GPU = torch.cuda.is_available()


def bench(fn, dev, warm=2, n=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        if dev == 'cuda':
            torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        if dev == 'cuda':
            torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    return min(ts), float(np.median(ts))


def mk(cfgname, aabb, dev, overrides=None, seed=20211202):
    torch.manual_seed(seed); np.random.seed(seed)
    cfg = load_cfg(cfgname); cfg.dataset.aabb = aabb
    for k, v in (overrides or {}).items():
        cfg.model[k] = v
    return cfg, FactorFields(cfg, dev)


def regress(cfgname, aabb, in_dim, out_dim, lr_small, lr_large, tag):
    res = {}
    for dev in (['cuda'] if HAS_GPU else []) + ['cpu']:
        cfg, m = mk(cfgname, aabb, dev)
        B = cfg.training.batch_size
        opt = torch.optim.Adam(m.get_optparam_groups(lr_small, lr_large), betas=(0.9, 0.99))
        g = torch.Generator().manual_seed(0)
        hi = torch.tensor(aabb[1][:in_dim]).float()
        if in_dim == 2:
            x = (torch.floor(torch.rand(B, 2, generator=g) * hi) + 0.5).to(dev)
        else:
            x = (torch.rand(B, 3, generator=g) * hi).to(dev)
        tgt = torch.rand(B, out_dim, generator=g).to(dev)

        def step():
            feats, _ = m.get_coding(x); y = m.linear_mat(feats)
            loss = torch.mean((y - tgt) ** 2); opt.zero_grad(); loss.backward(); opt.step()

        def fwd():
            with torch.no_grad():
                feats, _ = m.get_coding(x); m.linear_mat(feats)

        def coding():
            with torch.no_grad():
                m.get_coding(x)
        n = 20 if dev == 'cuda' else 3
        bs, ms = bench(step, dev, n=n); bf, mf = bench(fwd, dev, n=n); bc, mc = bench(coding, dev, n=n)
        res[dev] = {'B': B, 'step_ms_best': bs * 1e3, 'step_ms_med': ms * 1e3, 'fwd_ms_best': bf * 1e3,
                    'get_coding_ms_best': bc * 1e3, 'train_Mq_s': B / bs / 1e6, 'fwd_Mq_s': B / bf / 1e6,
                    'get_coding_Mq_s': B / bc / 1e6, 'n_params': m.n_parameters()}
        print(tag, dev, res[dev], flush=True)
    out[tag] = res


def synth_rays(n, seed, H=800, W=800, focal=1111.111, radius=4.0 / 1.5):
    """Blender-shaped pinhole rays: random cameras on a sphere looking at the origin, random pixels."""
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(n, 3, generator=g); c[:, 2] = c[:, 2].abs(); c = c / c.norm(dim=-1, keepdim=True) * radius
    fwd = -c / c.norm(dim=-1, keepdim=True)
    up = torch.tensor([0., 0., 1.]).expand_as(fwd)
    right = torch.cross(fwd, up, dim=-1); right = right / right.norm(dim=-1, keepdim=True)
    down = torch.cross(fwd, right, dim=-1)
    px = torch.rand(n, 2, generator=g) * torch.tensor([W, H]).float()
    d = fwd + right * ((px[:, :1] - W / 2) / focal) + down * ((px[:, 1:] - H / 2) / focal)
    d = d / d.norm(dim=-1, keepdim=True)
    return torch.cat([c, d], -1)


def synth_rgb(rays):
    """Analytic target: lambert-ish coloured sphere r=0.55 at origin on white background."""
    o, d = rays[:, :3], rays[:, 3:6]
    b = (o * d).sum(-1); cc = (o * o).sum(-1) - 0.55 ** 2
    disc = b * b - cc; hit = disc > 0
    t = -b - torch.sqrt(disc.clamp(min=0))
    p = o + d * t[:, None]; nrm = p / 0.55
    col = 0.5 + 0.5 * nrm
    return torch.where(hit[:, None], col, torch.ones_like(col))


def nerf(tag, overrides=None, aabb=None, train_iters=400):
    aabb = aabb or [[-1., -1., -1.], [1., 1., 1.]]
    res = {}
    n = 4096
    states = {}
    for dev in (['cuda'] if HAS_GPU else []) + ['cpu']:
        cfg, m = mk('nerf.yaml', aabb, dev, overrides)
        opt = torch.optim.Adam(m.get_optparam_groups(0.001, 0.02), betas=(0.9, 0.99))
        reso_cur = N_to_reso(128 ** 3, m.aabb)
        nS = min(cfg.renderer.max_samples, cal_n_samples(reso_cur, cfg.renderer.step_ratio))
        cnt = {}
        if dev == 'cpu' and 'cuda' in states:
            m.load_state_dict(states['cuda'])  # time CPU in the same (trained-ish) state

        def step(i=[0]):
            rays = synth_rays(n, 1000 + i[0]); tgt = synth_rgb(rays).to(dev); i[0] += 1
            rgb, depth, coef = R.render_ray(rays, m, chunk=n, N_samples=nS, white_bg=True, is_train=True, device=dev)
            cnt['valid'] = coef.shape[0]
            loss = torch.mean((rgb - tgt) ** 2); opt.zero_grad(); loss.backward(); opt.step()
            cnt['loss'] = loss
        r = {'n_rays': n, 'nSamples': nS, 'n_params': m.n_parameters()}
        if dev == 'cuda':
            b0, m0 = bench(step, dev, n=10)
            r['init_state'] = {'step_ms_best': b0 * 1e3, 'rays_s': n / b0, 'valid_frac': cnt['valid'] / n / nS,
                               'Mq_s': cnt['valid'] / b0 / 1e6}
            t = time.perf_counter()
            for _ in range(train_iters):
                step()
            torch.cuda.synchronize(); r['warm_train_iters'] = train_iters
            r['warm_train_s'] = time.perf_counter() - t
            r['loss_after'] = float(cnt['loss'])
            states['cuda'] = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
        b1, m1 = bench(step, dev, n=10 if dev == 'cuda' else 2, warm=1)
        # count app-mask survivors in the current state
        with torch.no_grad():
            rays = synth_rays(n, 7).to(dev)
            xyz, z, msk = m.sample_point(rays[:, :3], rays[:, 3:6], is_train=False, N_samples=nS)
            feats, _ = m.get_coding(xyz[msk]); feat = m.linear_mat(feats)
            sigma = torch.zeros(xyz.shape[:-1], device=dev); sigma[msk] = m.basis2density(feat[..., 0])
            dists = torch.cat((z[:, 1:] - z[:, :-1], torch.zeros_like(z[:, :1])), -1)
            from models.FactorFields import raw2alpha
            _, w, _ = raw2alpha(sigma, dists * cfg.renderer.distance_scale)
            app = (w > cfg.renderer.rayMarch_weight_thres) & msk
        r['trained_state'] = {'step_ms_best': b1 * 1e3, 'step_ms_med': m1 * 1e3, 'rays_s': n / b1,
                              'valid_frac': cnt['valid'] / n / nS, 'Mq_s': cnt['valid'] / b1 / 1e6,
                              'app_frac_of_all': float(app.float().mean()), 'loss': float(cnt['loss'])}
        # eval-style forward (no grad, no jitter), chunk 4096
        rays_e = synth_rays(n, 99)

        def ev():
            with torch.no_grad():
                R.render_ray(rays_e, m, chunk=n, N_samples=nS, white_bg=True, is_train=False, device=dev)
        be, me = bench(ev, dev, n=10 if dev == 'cuda' else 2, warm=1)
        r['eval_fwd'] = {'ms_best': be * 1e3, 'rays_s': n / be}
        res[dev] = r
        print(tag, dev, r, flush=True)
        if dev == 'cuda':
            states['model_cuda'] = m; states['cfg'] = cfg; states['nS'] = nS
        else:
            states['model_cpu'] = m
    # reference self-consistency CPU vs CUDA on identical weights/inputs (no jitter)
    if HAS_GPU:
        mc, mg, nS = states['model_cpu'], states['model_cuda'], states['nS']
        with torch.no_grad():
            rays = synth_rays(2048, 5)
            xc, zc, kc = mc.sample_point(rays[:, :3], rays[:, 3:6], is_train=False, N_samples=nS)
            xg, zg, kg = mg.sample_point(rays[:, :3].cuda(), rays[:, 3:6].cuda(), is_train=False, N_samples=nS)
            par = {'mask_mismatch': int((kc != kg.cpu()).sum()), 'mask_total': int(kc.numel()),
                   'xyz_bitexact': bool(torch.equal(xc, xg.cpu())), 'xyz_maxabs': float((xc - xg.cpu()).abs().max())}
            pts = xc[kc][:200000]
            fc, _ = mc.get_coding(pts); fg, _ = mg.get_coding(pts.cuda()); fg = fg.cpu()
            den = fc.abs().max()
            par['feat_maxabs'] = float((fc - fg).abs().max()); par['feat_absmax'] = float(den)
            par['feat_rel_to_max'] = float((fc - fg).abs().max() / den)
            hc = mc.linear_mat(fc); hg = mg.linear_mat(fg.cuda()).cpu()
            par['linear_mat_maxabs'] = float((hc - hg).abs().max()); par['linear_mat_absmax'] = float(hc.abs().max())
            rc, dc = R.render_ray(rays, mc, chunk=2048, N_samples=nS, white_bg=True, is_train=False, device='cpu')
            rg, dg = R.render_ray(rays, mg, chunk=2048, N_samples=nS, white_bg=True, is_train=False, device='cuda')
            par['rgb_maxabs'] = float((rc - rg.cpu()).abs().max()); par['depth_maxabs'] = float((dc - dg.cpu()).abs().max())
        res['cpu_vs_cuda_reference'] = par
        print(tag, 'cpu_vs_cuda', par, flush=True)
    out[tag] = res


if __name__ == '__main__':
    which = sys.argv[1:] or ['image', 'sdf', 'nerf', 'nerf_vm', 'nerf_cp']
    if 'image' in which:
        regress('image.yaml', [[0., 0.], [1024, 1024]], 2, 3, 0.002, 0.002, 'image_1024')
    if 'sdf' in which:
        regress('sdf.yaml', [[0., 0., 0.], [640, 640, 640]], 3, 1, 0.002, 0.02, 'sdf_640')
    if 'nerf' in which:
        nerf('nerf_grid')
    if 'nerf_vm' in which:
        nerf('nerf_vm', dict(coeff_type='vm', basis_type='vm'), train_iters=200)
    if 'nerf_cp' in which:
        nerf('nerf_cp', dict(coeff_type='vec', basis_type='cp', freq_bands=[1.] * 6, basis_resos=[512] * 6,
                             basis_dims=[32] * 6), train_iters=200)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(out, open('gpurun_out/ref_probe.json', 'w'), indent=1)
    print(json.dumps(out, indent=1))
