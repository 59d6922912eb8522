"""Host-side ray synthesis for a virtual observer (numpy): pose_spherical
(sunerf/train/coordinate_transformation.py:36-54) and get_rays (sunerf/data/ray_sampling.py:7-36) restated for a
regular helioprojective pixel grid, as SURVEY.md section 8d specifies for the synthetic workloads."""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

SOLRAD_M = 6.957e8
R_OBS = 1.495978707e11 / SOLRAD_M      # 1 AU in solar radii


def pose_spherical(theta: float, phi: float, radius: float, shift: Optional[Sequence[float]] = None) -> np.ndarray:
    """sunerf/train/coordinate_transformation.py:36-54 - 4x4 camera-to-world pose, float32 matrix products in the
    reference's order: flip @ rot_theta @ rot_phi @ trans_t, then the optional translation.  (The one host-side
    definition of the package.)"""
    t = np.eye(4, dtype=np.float32); t[2, 3] = radius
    rp = np.array([[1, 0, 0, 0], [0, np.cos(phi), -np.sin(phi), 0], [0, np.sin(phi), np.cos(phi), 0], [0, 0, 0, 1]], dtype=np.float32)
    rt = np.array([[np.cos(theta), 0, -np.sin(theta), 0], [0, 1, 0, 0], [np.sin(theta), 0, np.cos(theta), 0], [0, 0, 0, 1]], dtype=np.float32)
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
    c2w = flip @ (rt @ (rp @ t))
    if shift is not None:
        ts = np.eye(4, dtype=np.float32); ts[:3, 3] = np.asarray(shift, dtype=np.float32)
        c2w = ts @ c2w
    return c2w


def observer_rays(H: int, W: int, plate_arcsec: float, lat_deg: float, lon_deg: float,
                  distance: float = R_OBS) -> Tuple[np.ndarray, np.ndarray]:
    """rays_o, rays_d ([H*W,3] float32) for an H x W image with the given plate scale seen from (lat, lon, distance)."""
    c2w = pose_spherical(-np.deg2rad(lon_deg), np.deg2rad(lat_deg), distance)
    ii, jj = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
    rad = np.pi / 180 / 3600
    Tx, Ty = (jj - (W - 1) / 2) * plate_arcsec * rad, (ii - (H - 1) / 2) * plate_arcsec * rad
    dirs = np.stack([np.sin(Tx), -np.sin(Ty) * np.cos(Tx), -np.cos(Tx) * np.cos(Ty)], -1).astype(np.float32)
    rays_d = np.sum(dirs[..., None, :] * c2w[:3, :3], axis=-1).astype(np.float32).reshape(-1, 3)
    rays_o = np.broadcast_to(c2w[:3, -1], rays_d.shape).astype(np.float32)
    return rays_o, rays_d


def synthetic_rays(n: int, seed: int = 0, H: int = 256, W: int = 256, plate_arcsec: float = 9.4, t_days: float = 30.0,
                   n_views: int = 3, n_channels: int = 1) -> Dict[str, torch.Tensor]:
    """Shuffled rays from `n_views` random near-ecliptic observers + U(0,t_days) times + U(0,1) targets."""
    rng = np.random.default_rng(seed)
    os_, ds_ = [], []
    for _ in range(n_views):
        o, d = observer_rays(H, W, plate_arcsec, lat_deg=rng.uniform(-7, 7), lon_deg=rng.uniform(0, 360))
        os_.append(o); ds_.append(d)
    o, d = np.concatenate(os_), np.concatenate(ds_)
    sel = rng.permutation(o.shape[0])[:n]
    return {'rays_o': torch.from_numpy(o[sel].copy()), 'rays_d': torch.from_numpy(d[sel].copy()),
            'times': torch.from_numpy(rng.uniform(0, t_days, (n, 1)).astype(np.float32)),
            'target': torch.from_numpy(rng.uniform(0, 1, (n, n_channels)).astype(np.float32))}
