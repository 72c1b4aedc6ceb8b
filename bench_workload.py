"""Synthetic workload shared by both bench arms (ours and --impl reference): the nerf.yaml shapes
(configs/nerf.yaml: aabb +-1, coefficient grid 18x32^3, basis levels {4,4,4,2,2,2} x {25,39,54,69,84,99}^3,
linear_mat 18->64->32, renderModule 194->128->128->3, 128^3 render grid -> 443 train samples per ray, 4096 rays
per step), seeded numpy only — no dataset, no checkpoint, no GPU."""
import numpy as np

AABB = [[-1., -1., -1.], [1., 1., 1.]]
COEFF_RESO = 32
BASIS_DIMS = [4, 4, 4, 2, 2, 2]
BASIS_RESO = [25, 39, 54, 69, 84, 99]
# FactorFields.py:302-306: cfg freq_bands * (scene_reso / max(basis_reso) / max(freq_bands)) in fp32
FREQ_BANDS = (np.array([2., 3.2, 4.4, 5.6, 6.8, 8.], np.float32) * np.float32(768 / 99.0 / 8.0)).astype(np.float32).tolist()
N_SAMPLES = 443          # cal_n_samples([128]*3, 0.5)
BATCH = 4096
RCFG = dict(density_shift=-10.0, distance_scale=25.0, rayMarch_weight_thres=1e-3, view_pe=6, fea_pe=2)


def _uniform(rng, shape, fan_in):
    b = 1.0 / np.sqrt(fan_in)
    return rng.uniform(-b, b, shape).astype(np.float32)


def make_state(seed=0, blob_gain=16.0):
    """'Mid-training-like' synthetic weights in the reference's state_dict naming/layout: random-init MLPs and bases,
    plus a smooth density blob wired through coefficient channel 0 -> hidden unit 0 -> density feature, so that a
    realistic share of samples passes the weight threshold and reaches the appearance MLP."""
    rng = np.random.RandomState(seed)
    sd = {}
    g = (np.arange(COEFF_RESO) + 0.5) / COEFF_RESO * 2 - 1
    zz, yy, xx = np.meshgrid(g, g, g, indexing='ij')
    blob = np.zeros_like(xx)
    for c, s in [((0.0, 0.0, -0.1), 0.38), ((0.35, -0.2, 0.25), 0.22), ((-0.4, 0.3, 0.1), 0.2)]:
        blob += np.exp(-((xx - c[0]) ** 2 + (yy - c[1]) ** 2 + (zz - c[2]) ** 2) / (2 * s * s))
    coeff = np.ones((1, sum(BASIS_DIMS), COEFF_RESO, COEFF_RESO, COEFF_RESO), np.float32)
    coeff += 0.05 * rng.randn(*coeff.shape).astype(np.float32)
    coeff[0, 0] = 1.0 + blob_gain * np.clip(blob, 0, 1).astype(np.float32)
    sd['coeffs.0'] = coeff
    for i, (c, r) in enumerate(zip(BASIS_DIMS, BASIS_RESO)):
        b = (0.3 * rng.randn(1, c, r, r, r)).astype(np.float32)
        if i == 0:
            b[0, 0] = 1.0                      # feats[:, 0] == coefficient channel 0
        sd[f'basises.{i}'] = b
    W1, b1 = _uniform(rng, (64, 18), 18), _uniform(rng, (64,), 18)
    W1[0] = 0; W1[0, 0] = 1.0; b1[0] = -1.0    # hidden unit 0 = relu(coeff0 - 1) = gain * blob
    W2 = _uniform(rng, (32, 64), 64)
    W2[0] = 0; W2[0, 0] = 1.0                  # density feature = hidden unit 0
    sd['linear_mat.backbone.0.weight'], sd['linear_mat.backbone.0.bias'], sd['linear_mat.backbone.1.weight'] = W1, b1, W2
    sd['renderModule.mlp.0.weight'], sd['renderModule.mlp.0.bias'] = _uniform(rng, (128, 194), 194), _uniform(rng, (128,), 194)
    sd['renderModule.mlp.1.weight'], sd['renderModule.mlp.1.bias'] = _uniform(rng, (128, 128), 128), _uniform(rng, (128,), 128)
    sd['renderModule.mlp.2.weight'] = _uniform(rng, (3, 128), 128)
    return sd


def make_rays(n, seed=0, radius=4.0 / 1.5, wh=(800, 800), focal=1111.11):
    """Blender-shaped rays (dataLoader/blender.py:50-90 conventions), vectorised.  wh / focal / radius: image size, focal
    length and camera distance (the Tanks&Temples-shaped variants use 1920x1080)."""
    rng = np.random.RandomState(seed)
    th, ph = rng.uniform(0, 2 * np.pi, n), rng.uniform(0.1, 0.45 * np.pi, n)
    c = radius * np.stack([np.cos(th) * np.sin(ph), np.sin(th) * np.sin(ph), np.cos(ph)], -1)
    fwd = -c / np.linalg.norm(c, axis=-1, keepdims=True)
    right = np.cross(fwd, np.array([0, 0, 1.0]))
    right /= np.linalg.norm(right, axis=-1, keepdims=True)
    up = np.cross(right, fwd)
    px, py = rng.uniform(0, wh[0], n), rng.uniform(0, wh[1], n)
    d = fwd + ((px - wh[0] / 2) / focal)[:, None] * right + ((py - wh[1] / 2) / focal)[:, None] * up
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rays = np.concatenate([c, d], -1).astype(np.float32)
    target = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    jitter = rng.uniform(0, 1, n).astype(np.float32)
    return rays, target, jitter


def step_size():
    """update_renderParams at gridSize 128^3, aabb +-1, step_ratio 0.5 (FactorFields.py:693-697), fp32."""
    units = np.float32(2.0) / np.float32(127.0)
    return np.float32(np.float32(units) * np.float32(0.5))


def regress_state(cfg, shapes, seed=0):
    """Seeded synthetic state (reference state_dict names / layout) for the regression workloads of bench.py's CPU leg:
    `shapes` = ffb200.models.FactorFields.field_shapes(cfg, aabb).  Grid coefficient + grid bases + linear_mat."""
    rng = np.random.RandomState(seed)
    d = int(shapes['in_dim'])
    dims = list(cfg.model.basis_dims)
    width = sum(dims)
    sd = {'coeffs.0': (cfg.model.coef_init + 0.05 * rng.randn(1, width, *[int(r) for r in shapes['coeff_reso']])).astype(np.float32)}
    for i, (c, r) in enumerate(zip(dims, shapes['basis_reso'])):
        sd[f'basises.{i}'] = (0.3 * rng.randn(1, c, *([int(r)] * d))).astype(np.float32)
    L, H, out = int(cfg.model.num_layers), int(cfg.model.hidden_dim), int(cfg.model.out_dim)
    for l in range(L):
        fi, fo = (width if l == 0 else H), (out if l == L - 1 else H)
        sd[f'linear_mat.backbone.{l}.weight'] = _uniform(rng, (fo, fi), fi)
        if l != L - 1:
            sd[f'linear_mat.backbone.{l}.bias'] = _uniform(rng, (fo,), fi)
    return sd


# ---- BASELINE config 4: the -vm / -CP presets of README_FactorField.md:12-32 on Tanks&Temples-shaped rays ------------------------
TNT_AABB = (1.2 * np.array([[-1.5, -0.6, -1.8], [1.5, 0.9, 1.8]])).tolist()      # non-cubic box
TNT_N_SAMPLES = 498                                                              # cal_n_samples(N_to_reso(128^3, aabb), 0.5)
PRESETS = {
    # name: (model overrides as run_batch.py:43 passes them, Fdim, B_fwd, B_bwd, linear_mat flop/query)   [SURVEY 8d table]
    'nerf_vm': ({'coeff_type': 'vm', 'basis_type': 'vm'}, 54, 1524, 4104, 11008),
    'nerf_cp': ({'coeff_type': 'vec', 'basis_type': 'cp', 'freq_bands': [1.] * 6, 'basis_resos': [512] * 6, 'basis_dims': [32] * 6}, 192, 5388, 14592, 28672),
}


def tnt_rays(n, seed=0):
    return make_rays(n, seed, radius=4.2, wh=(1920, 1080), focal=1165.0)


def density_offset_(model, value=10.2):
    """Mid-training-like state for a freshly initialised model (ours or the reference's: same module names): hidden unit 63 of
    linear_mat becomes the constant 1 and feeds `value` into the density feature, so that sigma = softplus(f0 - 10) is of
    order 1 and a few percent of the samples of every ray pass rayMarch_weight_thres and reach the appearance MLP (at init
    density_shift = -10 shades nothing).  In place, no_grad."""
    import torch
    with torch.no_grad():
        lm = model.linear_mat.backbone
        lm[0].weight[63].zero_()
        lm[0].bias[63] = 1.0
        lm[-1].weight[0].mul_(0.1)
        lm[-1].weight[0, 63] = value
    return model
