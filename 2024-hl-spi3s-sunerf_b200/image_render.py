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
from .rays import R_OBS, pose_spherical  # noqa: F401  (re-exported: image_render.pose_spherical)


def _rad(x) -> float:
    """float radians, or an astropy-like quantity (the reference passes `lat.to_value(u.rad)`)."""
    return float(x.to_value('rad')) if hasattr(x, 'to_value') else float(x)


class ObserverRenderer:
    """Renders novel views of a trained rendering module.

        r = ObserverRenderer(rendering, resolution=(1024, 1024), plate_arcsec=2.4)
        out = r.render_observer_image(lat, lon, time, wl=[94, 171, 193, 211, 304, 335])   # dict of [H, W, ...] arrays
    """

    def __init__(self, rendering, resolution: Tuple[int, int], plate_arcsec: float, device=None, use_cuda_graph: bool = True):
        """use_cuda_graph: full batches are rendered by replaying ONE captured batch render (static input / output buffers)
        instead of ~15 eager launches through Python per batch - the eager loop left the GPU idle for 14 % of a 1024^2
        image (round-1 bench: 343 vs 398 Msamples/s).  The ragged last batch runs eagerly."""
        self.rendering = rendering
        self.use_cuda_graph = bool(use_cuda_graph)
        self._graph, self._g_key, self._g_in, self._g_out = None, None, None, None
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
        sampler = self.rendering.sampler
        S = sampler.t_vals.shape[1]
        for a in range(0, count, batch_size):
            b = min(a + batch_size, count)
            # the reference's stratified jitter: one torch.rand([n, S]) per rendered batch (sampling.py:97), here from
            # `t_rand_generator` when one is given (reproducible renders)
            t_rand = torch.rand((b - a, S), device=self.device, generator=t_rand_generator) if sampler.perturb else None
            wl_b = None if wl_t is None else wl_t[None, :].expand(b - a, wl_t.shape[0])
            if self.use_cuda_graph and b - a == batch_size and count >= 2 * batch_size:
                out = self._replay(rays_o[a:b], rays_d[a:b], times[a:b], wl_b, t_rand)
            else:
                args = [rays_o[a:b], rays_d[a:b], times[a:b]] + ([] if wl_b is None else [wl_b.contiguous()])
                out = self.rendering(*args, t_rand=t_rand)
            for k, v in out.items():
                if k not in outputs:           # the image is assembled in place on the device: no per-batch cat
                    outputs[k] = torch.empty((count,) + tuple(v.shape[1:]), device=self.device, dtype=v.dtype)
                outputs[k][a:b] = v
        shaped = {k: v.view(r1 - r0, W, *v.shape[1:]) for k, v in outputs.items()}
        return {k: v.cpu().numpy() for k, v in shaped.items()} if as_numpy else shaped

    def _replay(self, rays_o, rays_d, times, wl, t_rand):
        """One full batch through the captured graph.  The graph holds no weights: the tensor-core path's packed weight
        image is refreshed (outside the graph, only when the parameters changed) before every replay."""
        ins = {'rays_o': rays_o, 'rays_d': rays_d, 'times': times}
        if wl is not None:
            ins['wl'] = wl
        if t_rand is not None:
            ins['t_rand'] = t_rand
        key = tuple((k, tuple(v.shape)) for k, v in ins.items())
        for m in (self.rendering.coarse_model, self.rendering.fine_model):
            if getattr(m, 'precision', None) in ops.TC_MODES:
                ps = m.linear_params()
                m._packed_ptr(ps[0::2], ps[1::2])
        if key != self._g_key:
            self._g_in = {k: v.detach().to(self.device, torch.float32).contiguous().clone() for k, v in ins.items()}
            gi = self._g_in
            call = lambda: self.rendering(gi['rays_o'], gi['rays_d'], gi['times'], *([gi['wl']] if 'wl' in gi else []),
                                          t_rand=gi.get('t_rand'))
            call()                                     # warm-up: lazy allocations, one-time kernel attributes, cached scalars
            torch.cuda.synchronize(self.device)
            self._graph = torch.cuda.CUDAGraph()
            l0 = ops.launch_count()
            with torch.cuda.graph(self._graph):
                self._g_out = call()
            self._g_launches = ops.launch_count() - l0
            from . import _lib
            _lib.lib().snf_count_launches(-self._g_launches)     # recorded, not executed
            self._g_key = key
        for k, v in ins.items():
            self._g_in[k].copy_(v, non_blocking=True)
        self._graph.replay()
        from . import _lib
        _lib.lib().snf_count_launches(self._g_launches)
        return self._g_out

    def render_sharded(self, rank: int, world: int, *args, **kw):
        """This rank's block of rows (ragged allowed) and its row slice; the host stitches the blocks, no collective."""
        H = self.resolution[0] if kw.get('resolution') is None else int(kw['resolution'][0])
        rows = parallel.shard_rows(H, rank, world)
        return rows, self.render_observer_image(*args, rows=rows, **kw)


def stitch_rows(parts: Sequence[Tuple[slice, Dict[str, np.ndarray]]]) -> Dict[str, np.ndarray]:
    """Concatenate per-rank row blocks (as returned by ObserverRenderer.render_sharded) in row order."""
    parts = sorted(parts, key=lambda p: p[0].start)
    return {k: np.concatenate([p[1][k] for p in parts], axis=0) for k in parts[0][1]}
