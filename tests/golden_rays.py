"""Synthetic Blender-shaped rays (dataLoader/blender.py:50-90 conventions: 800x800 pinhole, focal 1111.11, camera
centres on the upper hemisphere at radius 4/1.5, unit directions).  Same generator as tests/golden/make_golden.py."""
import numpy as np


def blender_like_rays(R, seed, radius=4.0 / 1.5):
    rng = np.random.RandomState(seed)
    rays = np.zeros((R, 6), np.float32)
    for i in range(R):
        th, ph = rng.uniform(0, 2 * np.pi), rng.uniform(0.1, 0.45 * np.pi)
        c = radius * np.array([np.cos(th) * np.sin(ph), np.sin(th) * np.sin(ph), np.cos(ph)])
        fwd = -c / np.linalg.norm(c)
        up = np.array([0, 0, 1.0])
        right = np.cross(fwd, up); right /= np.linalg.norm(right)
        up2 = np.cross(right, fwd)
        px, py = rng.uniform(0, 800, 2)
        d = fwd + (px - 400) / 1111.11 * right + (py - 400) / 1111.11 * up2
        d /= np.linalg.norm(d)
        rays[i, :3], rays[i, 3:] = c, d
    return rays
