"""Rendering layer with the reference's class surface (sunerf/rendering/{base_tracing,emission,density_temperature}.py).

    Rendering(Rs_per_ds, sampling_config=None, hierarchical_sampling_config=None, model=NeRF, model_config=None, ...)
    forward(rays_o[N,3], rays_d[N,3], times[N,1], wavelengths[N,C]=None) -> dict with the 8 reference keys

Differentiable through torch.autograd via analytic-backward Functions over the CUDA kernels; the training fast
path (trainer.py) calls the same kernels without autograd.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch import nn

from . import ops
from ._lib import SnfError
from .model import NeRF
from .sampling import HierarchicalSampler, SphericalSampler, StratifiedSampler

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')


def load_aia_response(aia_exp_time: float = 2.9):
    """The 7 x 101 AIA temperature-response table the reference reads from sunerf/data/aia_temp_resp.genx
    (density_temperature.py:131-146), shipped here as data/aia_response.npz (float32(TRESP*2.9), logT grid)."""
    d = np.load(os.path.join(_DATA, 'aia_response.npz'))
    if abs(aia_exp_time - 2.9) > 1e-12:
        raise SnfError('the shipped response table is pre-multiplied by the reference default aia_exp_time=2.9')
    assert tuple(d['channels']) == ops.AIA_CHANNELS
    return torch.from_numpy(d['logT'].copy()), torch.from_numpy(d['table'].copy())


# ------------------------------------------------------------------------------------------ autograd glue
class _EmissionComposite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, z, rays_d):
        image, weights, absorption = ops.composite_emission_fwd(raw, z, rays_d)
        ctx.save_for_backward(raw, z, rays_d)
        ctx.mark_non_differentiable(weights)
        return image, weights, absorption

    @staticmethod
    def backward(ctx, g_image, _g_weights, g_abs):
        raw, z, rays_d = ctx.saved_tensors
        g_raw = ops.composite_emission_bwd(raw, z, rays_d, g_image.reshape(-1), g_abs)
        return g_raw, None, None


class _DTComposite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inferences, z, wavelengths, log_abs, vol_c, table_x, table_y, F):
        image, weights, regq = ops.composite_dt_fwd(inferences, z, wavelengths, log_abs, vol_c, table_x, table_y, F)
        ctx.save_for_backward(inferences, z, wavelengths, log_abs, vol_c, table_x, table_y)
        ctx.F = F
        ctx.mark_non_differentiable(weights)
        return image, weights, regq

    @staticmethod
    def backward(ctx, g_image, _g_weights, g_regq):
        inferences, z, wavelengths, log_abs, vol_c, table_x, table_y = ctx.saved_tensors
        g_inf, g_la, g_vc = ops.composite_dt_bwd(inferences, z, wavelengths, log_abs, vol_c, table_x, table_y, ctx.F,
                                                 g_image, g_regq)
        return g_inf, None, None, g_la, g_vc.reshape(vol_c.shape), None, None, None


class _Epilogue(torch.autograd.Function):
    """height_map, absorption_map, regularization from the fine pass; differentiable w.r.t. q only (as the
    reference's loss only reaches q through `regularization`)."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, z_comb, weights, q, r0, kind):
        hm, am, reg, dq = ops.render_epilogue(rays_o, rays_d, z_comb, weights, q, r0, kind, grad_scale=1.0, want_gq=True)
        ctx.save_for_backward(dq)
        ctx.mark_non_differentiable(hm, am)
        return hm, am, reg

    @staticmethod
    def backward(ctx, _g_hm, _g_am, g_reg):
        (dq,) = ctx.saved_tensors
        return None, None, None, None, g_reg * dq, None, None


# ------------------------------------------------------------------------------------------ rendering classes
class SuNeRFRendering(nn.Module):
    """base_tracing.py:8-132"""

    kind = 0           # 0 emission-style regulariser, 1 density-temperature
    reg_radius = 1.2

    def __init__(self, Rs_per_ds, sampling_config=None, hierarchical_sampling_config=None, model=NeRF, model_config=None):
        super().__init__()
        self.Rs_per_ds = Rs_per_ds
        hierarchical_sampling_config = {'type': 'hierarchical'} if hierarchical_sampling_config is None \
            else dict(hierarchical_sampling_config)
        sampling_config = {'type': 'stratified'} if sampling_config is None else dict(sampling_config)
        model_config = {} if model_config is None else model_config
        sampling_type = sampling_config.pop('type')
        if sampling_type == 'stratified':
            self.sampler = StratifiedSampler(Rs_per_ds=Rs_per_ds, **sampling_config)
        elif sampling_type == 'spherical':
            self.sampler = SphericalSampler(Rs_per_ds=Rs_per_ds, **sampling_config)
        else:
            raise ValueError(f'Unknown sampling type {sampling_type}')
        hierarchical_sampling_type = hierarchical_sampling_config.pop('type')
        if hierarchical_sampling_type == 'hierarchical':
            self.sampler_hierarchical = HierarchicalSampler(**hierarchical_sampling_config)
        else:
            raise ValueError(f'Unknown sampling type {hierarchical_sampling_type}')
        self.coarse_model = model(**model_config)
        self.fine_model = model(**model_config)

    def forward(self, rays_o, rays_d, times, wavelengths=None, t_rand=None):
        """base_tracing.py:46-111.  `t_rand` (optional, not in the reference signature) injects the stratified
        jitter for cross-device parity tests; by default it is drawn with torch.rand exactly as the reference does."""
        if not rays_o.is_cuda:
            raise SnfError('sunerf_b200 renders on CUDA only (no CPU fallback)')
        rays_o, rays_d, times = rays_o.float().contiguous(), rays_d.float().contiguous(), times.float().contiguous()
        z_vals, _ = self.sampler.sample_z(rays_o, rays_d, t_rand=t_rand)
        query = ops.make_query(rays_o, rays_d, z_vals, times)
        coarse_out = self._render(self.coarse_model, query, rays_d, rays_o, z_vals, wavelengths)
        outputs = {'z_vals_stratified': z_vals, 'coarse_image': coarse_out['image']}

        z_hierarch, z_comb = self.sampler_hierarchical.resample(z_vals, coarse_out['weights'])
        query_f = ops.make_query(rays_o, rays_d, z_comb, times)
        fine_out = self._render(self.fine_model, query_f, rays_d, rays_o, z_comb, wavelengths)

        outputs['z_vals_hierarchical'] = z_hierarch
        outputs['fine_image'] = fine_out['image']
        hm, am, reg = _Epilogue.apply(rays_o, rays_d, z_comb, fine_out['weights'], fine_out['regularizing_quantity'],
                                      self.reg_radius / self.Rs_per_ds, self.kind)
        outputs['image'] = fine_out['image']
        outputs['height_map'] = hm
        outputs['absorption_map'] = am
        outputs['regularization'] = reg
        return outputs

    def forward_points(self, query_points):
        return self.fine_model(query_points.reshape(-1, 4))

    def _render(self, model, query_points, rays_d, rays_o, z_vals, wavelengths=None):
        shape = query_points.shape[:-1]
        raw = model(query_points.reshape(-1, 4))
        raw = raw['inferences'] if isinstance(raw, dict) else raw
        raw = raw.reshape(*shape, raw.shape[-1])
        return self.raw2outputs(raw=raw, z_vals=z_vals, rays_d=rays_d, rays_o=rays_o, query_points=query_points)

    def raw2outputs(self, **kwargs):
        raise NotImplementedError("This method should be implemented in a subclass")


class EmissionRadiativeTransfer(SuNeRFRendering):
    """emission.py:6-54"""

    kind = 0
    reg_radius = 1.2

    def __init__(self, model_config=None, **kwargs):
        model_config = {} if model_config is None else dict(model_config)
        model_config.update({'d_input': 4, 'd_output': 2})
        super().__init__(model_config=model_config, **kwargs)

    def raw2outputs(self, raw, z_vals, rays_d, **kwargs):
        image, weights, absorption = _EmissionComposite.apply(raw, z_vals, rays_d)
        return {'image': image, 'weights': weights, 'regularizing_quantity': absorption}


class DensityTemperatureRadiativeTransfer(SuNeRFRendering):
    """density_temperature.py:78-274"""

    kind = 1
    reg_radius = 1.25

    def __init__(self, model_config=None, device=None, aia_exp_time=2.9, pixel_intensity_factor=1e10, **kwargs):
        model_config = {} if model_config is None else model_config
        super().__init__(model_config=model_config, **kwargs)
        self.device = device
        self.pixel_intensity_factor = pixel_intensity_factor
        tx, ty = load_aia_response(aia_exp_time)
        # not persistent: the reference keeps its interpolators outside the state_dict too
        self.register_buffer('_table_x', tx, persistent=False)
        self.register_buffer('_table_y', ty, persistent=False)

    def _render(self, model, query_points, rays_d, rays_o, z_vals, wavelengths=None):
        if wavelengths is None:
            raise SnfError('DensityTemperatureRadiativeTransfer needs wavelengths[N,C]')
        shape = query_points.shape[:-1]
        state = model.forward(query_points.reshape(-1, 4))
        inf = state['inferences'].reshape(*shape, 2)
        return self.raw2outputs(inferences=inf, log_abs=state['log_abs'], vol_c=state['vol_c'], z_vals=z_vals,
                                rays_d=rays_d, wavelengths=wavelengths)

    def raw2outputs(self, inferences, log_abs, vol_c, z_vals, rays_d, wavelengths):
        la = torch.stack([log_abs[str(c)] for c in ops.AIA_CHANNELS])
        image, weights, regq = _DTComposite.apply(inferences, z_vals, wavelengths.float().contiguous(), la, vol_c,
                                                  self._table_x, self._table_y, float(self.pixel_intensity_factor))
        return {'image': image, 'weights': weights, 'regularizing_quantity': regq}

    def regularization(self, distance, regularizing_quantity):
        return torch.relu(distance - 1.25 / self.Rs_per_ds) * torch.relu(regularizing_quantity)
