"""Synthetic ray generators shared by tests/golden/make_golden.py (which records the reference's outputs on them) and
the GPU tests: llff-style NDC rays and 360-style inside-out rays.  numpy only."""
import numpy as np


def ndc_like_rays(R, seed):
    """Rays shaped like dataLoader/ray_utils.py ndc_rays_blender output (llff): origins on the z=-1 plane,
    directions with z = 2 (not unit length, so the |d| scaling of FactorFields.py:854-856 matters)."""
    rng = np.random.RandomState(seed)
    rays = np.zeros((R, 6), np.float32)
    rays[:, 0:2] = rng.uniform(-1.2, 1.2, (R, 2))
    rays[:, 2] = -1.0
    rays[:, 3:5] = rng.uniform(-0.6, 0.6, (R, 2))
    rays[:, 5] = 2.0
    return rays


def inside_out_rays(R, seed):
    """360-style rays: camera centres inside the unit cube, unit directions all around (the far samples leave the
    cube and get contracted, FactorFields.py:625-631)."""
    rng = np.random.RandomState(seed)
    rays = np.zeros((R, 6), np.float32)
    rays[:, :3] = rng.uniform(-0.6, 0.6, (R, 3))
    d = rng.normal(size=(R, 3))
    rays[:, 3:] = d / np.linalg.norm(d, axis=1, keepdims=True)
    return rays
