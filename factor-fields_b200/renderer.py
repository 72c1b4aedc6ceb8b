"""Ray-chunk renderer with the signature of the reference's renderer.render_ray (renderer.py:8-27): the host
boundary of the hot path.  Rays may live on the host (pinned or pageable); each chunk is copied to the device
here, exactly where the reference does its `.to(device)`."""
import torch


def render_ray(rays, factor_fields, chunk=4096, N_samples=-1, ndc_ray=False, white_bg=True, is_train=False, device='cuda'):
    rgbs, depth_maps, coeffs = [], [], []
    n_rays = rays.shape[0]
    for start in range(0, n_rays, chunk):
        rays_chunk = rays[start:start + chunk].to(device, non_blocking=True)
        rgb_map, depth_map, coeff = factor_fields(rays_chunk, is_train=is_train, white_bg=white_bg, ndc_ray=ndc_ray,
                                                  N_samples=N_samples)
        rgbs.append(rgb_map)
        depth_maps.append(depth_map)
        if is_train:
            coeffs.append(coeff)
    if is_train:
        return torch.cat(rgbs), torch.cat(depth_maps), torch.cat(coeffs)
    return torch.cat(rgbs), torch.cat(depth_maps)
