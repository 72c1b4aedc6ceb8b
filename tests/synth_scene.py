"""Analytic test scene for training-parity runs: a shaded sphere in front of a white background, seen by
Blender-shaped cameras (tests/golden_rays.py).  Target colour of a ray = 0.5 + 0.5 * normal at the first hit."""
import numpy as np

from tests.golden_rays import blender_like_rays


def sphere_scene(n_rays, seed, radius=0.55):
    rays = blender_like_rays(n_rays, seed)
    o, d = rays[:, :3].astype(np.float64), rays[:, 3:].astype(np.float64)
    b = (o * d).sum(-1)
    c = (o * o).sum(-1) - radius * radius
    disc = b * b - c
    hit = disc > 0
    t = -b - np.sqrt(np.where(hit, disc, 0.0))
    p = o + d * t[:, None]
    nrm = p / np.maximum(np.linalg.norm(p, axis=-1, keepdims=True), 1e-9)
    rgb = np.where(hit[:, None], 0.5 + 0.5 * nrm, 1.0).astype(np.float32)
    return rays, rgb


def synth_image(H, W, channels=3, seed=0, n_waves=24):
    """Band-limited random field in [0, 1] (sum of random sinusoids + a smooth gradient) standing in for the image of
    scripts/2D_regression.ipynb; -> (coords [H*W, 2] pixel centres (x, y) + 0.5 as dataLoader/image.py:44-45, img [H*W, C])."""
    rng = np.random.RandomState(seed)
    y, x = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
    img = np.zeros((H, W, channels))
    for c in range(channels):
        for _ in range(n_waves):
            fx, fy = rng.uniform(-0.12, 0.12, 2)
            img[..., c] += rng.uniform(0.2, 1.0) * np.sin(2 * np.pi * (fx * x + fy * y) + rng.uniform(0, 2 * np.pi))
        img[..., c] = img[..., c] / (2.5 * np.sqrt(n_waves) * 0.6) + 0.5 + 0.2 * (x / W - 0.5) * (1 if c % 2 else -1)
    img = np.clip(img, 0.0, 1.0).astype(np.float32)
    coords = (np.stack([x, y], -1).reshape(-1, 2) + 0.5).astype(np.float32)
    return coords, img.reshape(H * W, channels)


def synth_sdf(n, extent, seed=0):
    """Analytic signed-distance samples standing in for the mesh SDF files of scripts/sdf_regression.ipynb: points uniform
    in [0, extent]^3 (the reference maps its [-1,1]^3 samples there, dataLoader/sdf.py:21-28), target = SDF of the union
    of two spheres and a torus in the normalised cube."""
    rng = np.random.RandomState(seed)
    p = rng.uniform(-1, 1, (n, 3))
    s1 = np.linalg.norm(p - np.array([0.25, 0.1, 0.0]), axis=1) - 0.45
    s2 = np.linalg.norm(p - np.array([-0.35, -0.2, 0.2]), axis=1) - 0.3
    q = np.stack([np.linalg.norm(p[:, :2], axis=1) - 0.6, p[:, 2] + 0.4], 1)
    tor = np.linalg.norm(q, axis=1) - 0.12
    sdf = np.minimum(np.minimum(s1, s2), tor)
    coords = ((p + 1) * 0.5 * extent).astype(np.float32)
    return coords, sdf.astype(np.float32)[:, None]


def synth_image_set(n_img, H, W, seed=0):
    """n_img small images (a shared band-limited pattern + a per-image blob) with the coordinates of
    dataLoader/image_set.py: (x + 0.5, y + 0.5, image + 0.5)."""
    rng = np.random.RandomState(seed)
    base_c, base = synth_image(H, W, 3, seed + 1, n_waves=10)
    coords, imgs = [], []
    for i in range(n_img):
        cx, cy, r = rng.uniform(0.2 * W, 0.8 * W), rng.uniform(0.2 * H, 0.8 * H), rng.uniform(0.1, 0.3) * W
        d2 = (base_c[:, 0] - cx) ** 2 + (base_c[:, 1] - cy) ** 2
        blob = np.exp(-d2 / (r * r))[:, None] * rng.uniform(0.2, 0.9, 3)[None]
        imgs.append(np.clip(0.5 * base + 0.5 * blob, 0, 1).astype(np.float32))
        coords.append(np.concatenate([base_c, np.full((H * W, 1), i + 0.5, np.float32)], 1))
    return np.concatenate(coords).astype(np.float32), np.concatenate(imgs).astype(np.float32)
