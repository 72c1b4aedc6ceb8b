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
