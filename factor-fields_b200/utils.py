"""Host-side helpers that decide every grid shape and sample count (reference utils.py:53-84); behaviour
(including fp32 rounding) must match the reference because grid shapes and nSamples derive from them."""
import math

import numpy as np
import torch


def N_to_reso(n_voxels, bbox):
    """utils.py:53-57 — per-axis resolution of a box holding ~n_voxels cubic voxels."""
    lo, hi = bbox
    extent = hi - lo
    voxel = (extent.prod() / n_voxels).pow(1 / len(lo))
    return torch.round(extent / voxel).long().tolist()


def N_to_vm_reso(n_voxels, bbox):
    """utils.py:59-67 — plane resolutions so that the three planes together hold ~n_voxels texels."""
    lo, hi = bbox
    extent = hi - lo
    if len(extent) != 3:
        raise AssertionError('vm factors need a 3-D box')
    reso = extent / (extent.prod() / n_voxels).pow(1 / len(lo))
    n_mat = reso[0] * reso[1] + reso[0] * reso[2] + reso[1] * reso[2]
    return torch.round(reso * math.sqrt(n_voxels / n_mat)).long().tolist()


def cal_n_samples(reso, step_ratio=0.5):
    """utils.py:69-70"""
    return int(np.linalg.norm(reso) / step_ratio)


class SimpleSampler:
    """utils.py:72-84 — epoch-wise numpy permutation of ray indices (consumes numpy's global RNG exactly like the
    reference so that both implementations see identical batches)."""

    def __init__(self, total, batch):
        self.total, self.batch = total, batch
        self.curr, self.ids = total, None

    def nextids(self):
        self.curr += self.batch
        if self.curr + self.batch > self.total:
            self.ids = torch.LongTensor(np.random.permutation(self.total))
            self.curr = 0
        return self.ids[self.curr:self.curr + self.batch]


def mse2psnr(mse):
    """renderer.py:60 / train_per_scene.py:166: -10 ln(mse) / ln(10)"""
    return -10.0 * math.log(mse) / math.log(10.0)


def remove_small_objects(mask, min_size, connectivity=1):
    """Replacement for skimage.morphology.remove_small_objects (FactorFields.py:774), which is not installed:
    drop connected components (face connectivity) with fewer than min_size voxels."""
    from scipy import ndimage
    structure = ndimage.generate_binary_structure(mask.ndim, connectivity)
    labels, n = ndimage.label(mask, structure=structure)
    if n == 0:
        return mask.copy()
    sizes = np.bincount(labels.ravel())
    too_small = sizes < min_size
    too_small[0] = False
    out = mask.copy()
    out[too_small[labels]] = False
    return out
