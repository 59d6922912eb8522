"""Full-image render driver: observer pose -> rays generated on the device -> batched
render -> image assembled on the device.  Replaces `SuNeRFLoader/ModelLoader.render_observer_image`
(sunerf/evaluation/loader.py:63-110, 151-242), i.e. the host numpy `get_rays` (sunerf/data/ray_sampling.py:7-36),
the per-batch `nn.DataParallel` scatter/gather (:37-39, 143-144) and the `ThreadPoolExecutor` + `torch.cat` stitching
(:226-242).  Same call shape and the same result dict (one [H, W, ...] array per key of the rendering's output).

The pixel grid is the regular helioprojective plate-scale grid; the reference takes it from a
sunpy map's WCS (third-party, not in this image).  Rows can be sharded over ranks (`rows=` / `parallel.shard_rows`):
rendering is ray-parallel and uses no collective.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops, parallel
from ._lib import SnfError
from .rays import R_OBS


def _rad(x) -> float:
    """float radians, or an astropy-like quantity (the reference passes `lat.to_value(u.rad)`)."""
    return float(x.to_value('rad')) if hasattr(x, 'to_value') else float(x)


def pose_spherical(theta: float, phi: float, radius: float, shift: Optional[Sequence[float]] = None) -> np.ndarray:
    """sunerf/train/coordinate_transformation.py:36-54 - 4x4 camera-to-world pose, float32 matrix products in the
    reference's order: flip @ rot_theta @ rot_phi @ trans_t, then the optional translation."""
    t = np.eye(4, dtype=np.float32); t[2, 3] = radius
    rp = np.array([[1, 0, 0, 0], [0, np.cos(phi), -np.sin(phi), 0], [0, np.sin(phi), np.cos(phi), 0], [0, 0, 0, 1]], dtype=np.float32)
    rt = np.array([[np.cos(theta), 0, -np.sin(theta), 0], [0, 1, 0, 0], [np.sin(theta), 0, np.cos(theta), 0], [0, 0, 0, 1]], dtype=np.float32)
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
    c2w = flip @ (rt @ (rp @ t))
    if shift is not None:
        ts = np.eye(4, dtype=np.float32); ts[:3, 3] = np.asarray(shift, dtype=np.float32)
        c2w = ts @ c2w
    return c2w


class ObserverRenderer:
    """Renders novel views of a trained rendering module.

        r = ObserverRenderer(rendering, resolution=(1024, 1024), plate_arcsec=2.4)
        out = r.render_observer_image(lat, lon, time, wl=[94, 171, 193, 211, 304, 335])   # dict of [H, W, ...] arrays
    """

    def __init__(self, rendering, resolution: Tuple[int, int], plate_arcsec: float, device=None):
        self.rendering = rendering
        self.resolution = tuple(int(v) for v in resolution)
        self.plate_arcsec = float(plate_arcsec)
        self.device = torch.device(device) if device is not None else next(rendering.parameters()).device
        if self.device.type != 'cuda':
            raise SnfError('ObserverRenderer needs the rendering module on a CUDA device')

    @torch.no_grad()
    def render_observer_image(self, lat, lon, time: float, distance: float = R_OBS, wl=None,
                              center: Optional[Sequence[float]] = None, resolution: Optional[Tuple[int, int]] = None,
                              batch_size: int = 4096, rows: Optional[slice] = None, as_numpy: bool = True,
                              t_rand_generator: Optional[torch.Generator] = None) -> Dict[str, object]:
        """lat, lon: radians (or astropy quantities); time: normalised time; distance: solar radii; wl: wavelengths of
        the channels to render (density-temperature model) or None (emission model); resolution: (H, W) override at
        the same field of view (the reference resamples its map); rows: render only this slice of image rows."""
        H0, W0 = self.resolution
        H, W = (H0, W0) if resolution is None else (int(resolution[0]), int(resolution[1]))
        plate = self.plate_arcsec * (W0 / W)          # Map.resample keeps the field of view
        c2w = pose_spherical(-_rad(lon), _rad(lat), float(distance.to_value('solRad')) if hasattr(distance, 'to_value') else float(distance), center)
        r0, r1, _ = (rows or slice(0, H)).indices(H)
        first, count = r0 * W, (r1 - r0) * W
        center_pixel = ((W - 1) / 2, (H - 1) / 2)
        rays_o, rays_d = ops.image_rays(c2w, H, W, plate, self.device, first, count, center_pixel)
        times = torch.full((count, 1), float(time), device=self.device, dtype=torch.float32)
        wl_t = None
        if wl is not None:
            wl_t = torch.as_tensor(np.asarray(wl, dtype=np.float32), device=self.device)
        outputs: Dict[str, torch.Tensor] = {}
        for a in range(0, count, batch_size):
            b = min(a + batch_size, count)
            args = [rays_o[a:b], rays_d[a:b], times[a:b]]
            if wl_t is not None:
                args.append(wl_t[None, :].expand(b - a, wl_t.shape[0]).contiguous())
            out = self.rendering(*args)
            for k, v in out.items():
                if k not in outputs:           # the image is assembled in place on the device: no per-batch cat
                    outputs[k] = torch.empty((count,) + tuple(v.shape[1:]), device=self.device, dtype=v.dtype)
                outputs[k][a:b] = v
        shaped = {k: v.view(r1 - r0, W, *v.shape[1:]) for k, v in outputs.items()}
        return {k: v.cpu().numpy() for k, v in shaped.items()} if as_numpy else shaped

    def render_sharded(self, rank: int, world: int, *args, **kw):
        """This rank's block of rows (ragged allowed) and its row slice; the host stitches the blocks, no collective."""
        H = self.resolution[0] if kw.get('resolution') is None else int(kw['resolution'][0])
        rows = parallel.shard_rows(H, rank, world)
        return rows, self.render_observer_image(*args, rows=rows, **kw)


def stitch_rows(parts: Sequence[Tuple[slice, Dict[str, np.ndarray]]]) -> Dict[str, np.ndarray]:
    """Concatenate per-rank row blocks (as returned by ObserverRenderer.render_sharded) in row order."""
    parts = sorted(parts, key=lambda p: p[0].start)
    return {k: np.concatenate([p[1][k] for p in parts], axis=0) for k in parts[0][1]}
