"""Samplers with the reference's module surface (sunerf/train/sampling.py) over the K1/K2 kernels."""
from __future__ import annotations

import torch

from . import ops


class StratifiedSampler(torch.nn.Module):
    """sampling.py:56-102.  Buffers distance / solar_R / t_vals[1,S] are state_dict keys of the reference."""

    def __init__(self, Rs_per_ds, distance=1.3, n_samples=64, perturb=True):
        super().__init__()
        self.perturb = perturb
        self.register_buffer('distance', torch.tensor(distance / Rs_per_ds, dtype=torch.float32))
        self.register_buffer('solar_R', torch.tensor(1 / Rs_per_ds, dtype=torch.float32))
        self.register_buffer('t_vals', torch.linspace(0., 1., n_samples)[None].to(torch.float32))
        # plain-float copies so the launch needs no device->host read
        self._distance = float(torch.tensor(distance / Rs_per_ds, dtype=torch.float32))
        self._solar_R = float(torch.tensor(1 / Rs_per_ds, dtype=torch.float32))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._distance, self._solar_R = float(self.distance.cpu()), float(self.solar_R.cpu())

    def sample_z(self, rays_o, rays_d, t_rand=None, want_points=False):
        """t_rand: the torch.rand([N,S]) draw of sampling.py:97; drawn here (same call, shape, device) if None."""
        if self.perturb and t_rand is None:
            t_rand = torch.rand((rays_o.shape[0], self.t_vals.shape[1]), device=rays_o.device)
        if not self.perturb:
            t_rand = None
        return ops.stratified_sample(rays_o, rays_d, self.t_vals, t_rand, self._distance, self._solar_R, want_points,
                                     spherical=getattr(self, '_kind', 'stratified') == 'spherical')

    def forward(self, rays_o: torch.Tensor, rays_d: torch.Tensor):
        z, pts = self.sample_z(rays_o, rays_d, want_points=True)
        return {'points': pts, 'z_vals': z}


class SphericalSampler(StratifiedSampler):
    """sampling.py:4-54: bins between the ray's entry and exit of the sphere of radius `distance` (default 2 R_sun), the far end
    clipped at the solar surface; rays that miss the sphere get NaN rows, as in the reference.  Same buffers, same jitter
    draw, same kernel as the stratified sampler with the other pair of bin ends."""

    def __init__(self, Rs_per_ds, distance=2.0, n_samples=64, perturb=True):
        super().__init__(Rs_per_ds, distance=distance, n_samples=n_samples, perturb=perturb)
        self._kind = 'spherical'


class HierarchicalSampler(torch.nn.Module):
    """sampling.py:104-169.  perturb=False (the reference default): u = linspace(0, 1, n).  perturb=True (:144-146): one row
    of torch.rand draws per ray, drawn here with the reference's call (shape [N, n], the rays' device) or passed as `u`."""

    def __init__(self, n_samples=128, perturb=False):
        super().__init__()
        self.n_samples, self.perturb = n_samples, perturb
        self._u = None

    def u(self, device):
        if self._u is None or self._u.device != device:
            self._u = torch.linspace(0., 1., self.n_samples).to(device)   # host linspace == the reference's values
        return self._u

    def resample(self, z_vals, weights, u=None):
        if self.perturb or u is not None:
            if u is None:
                u = torch.rand(list(z_vals.shape[:-1]) + [self.n_samples], device=z_vals.device)     # sampling.py:145
            new_z, z_comb, _, _ = ops.hier_resample(z_vals, weights.detach(), u, per_ray_u=True)
        else:
            new_z, z_comb, _, _ = ops.hier_resample(z_vals, weights.detach(), self.u(z_vals.device))
        return new_z, z_comb

    def forward(self, rays_o, rays_d, z_vals, weights):
        new_z, z_comb = self.resample(z_vals, weights)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_comb[..., :, None]
        return {'points': pts, 'z_vals': z_comb, 'new_z_samples': new_z}
