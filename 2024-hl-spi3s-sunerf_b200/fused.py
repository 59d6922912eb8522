"""The whole render chain as one C call per direction: `snf_render_fused_fwd` / `snf_render_fused_bwd`
(include/sunerf_b200.h).  This is the binding a non-Python host would write; here it is exercised from ctypes so that the
parity tests can hold it against the stage-by-stage path of `rendering.py` / `trainer.py` (same kernels, same order:
SuNeRFRendering.forward, sunerf/rendering/base_tracing.py:46-111)."""
from __future__ import annotations

import ctypes
from typing import Dict

import torch

from . import _lib, ops
from ._lib import SnfError
from .rendering import DensityTemperatureRadiativeTransfer, SuNeRFRendering


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr_array(ts) -> ctypes.Array:
    return (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


class FusedRender:
    """Descriptor + workspace for one rendering module and one batch size.

    fwd = FusedRender(rendering, n_rays, train=True)
    out = fwd.forward(rays_o, rays_d, times, wavelengths=None, t_rand=None, reg_grad_scale=...)
    grads = fwd.backward(g_coarse_image, g_fine_image)      # {'coarse_model': (gW[9], gB[9]), 'fine_model': ...}
    """

    def __init__(self, rendering: SuNeRFRendering, n_rays: int, train: bool = False):
        self.r, self.N, self.train = rendering, int(n_rays), bool(train)
        self.dt = isinstance(rendering, DensityTemperatureRadiativeTransfer)
        dev = next(rendering.parameters()).device
        if dev.type != 'cuda':
            raise SnfError('FusedRender needs the rendering module on a CUDA device (no CPU fallback)')
        self.dev = dev
        cm, fm = rendering.coarse_model, rendering.fine_model
        if cm.precision != fm.precision:
            raise SnfError('coarse and fine model must use the same precision mode')
        self.mode = ops.MLP_MODES[cm.precision]
        d = _lib.RenderDesc()
        d.kind, d.mode = (1 if self.dt else 0), self.mode
        ps_c, ps_f = cm.linear_params(), fm.linear_params()
        self.w_c, self.b_c, self.w_f, self.b_f = ps_c[0::2], ps_c[1::2], ps_f[0::2], ps_f[1::2]
        d.n_hidden, d.d_filter = len(self.w_c) - 1, self.w_c[0].shape[0]
        d.out_offset0, d.out_offset1 = [float(v) for v in cm._out_offsets()]
        s = rendering.sampler
        if getattr(s, '_kind', 'stratified') != 'stratified' or rendering.sampler_hierarchical.perturb:
            raise SnfError('the one-call C entry renders the default sampler pair (stratified + unperturbed hierarchical); '
                           'SphericalSampler / HierarchicalSampler(perturb=True) run through the staged classes')
        self.t_vals = s.t_vals.reshape(-1).contiguous()
        self.u = rendering.sampler_hierarchical.u(dev)
        d.t_vals, d.u, d.S, d.n_new = self.t_vals.data_ptr(), self.u.data_ptr(), self.t_vals.numel(), self.u.numel()
        d.distance, d.solar_R = s._distance, s._solar_R
        d.reg_radius = rendering.reg_radius / rendering.Rs_per_ds
        self.perturb = bool(s.perturb)
        self.C = 1
        self.desc = d
        self._ws = None

    # -- per-call refresh of what may have changed since the last call (weights re-packed, log_abs values)
    def _refresh(self, C: int):
        d, r = self.desc, self.r
        # host arrays of device pointers (the parameters may have been re-homed, e.g. into RayTrainer's flat buffer)
        self._arrs = [_ptr_array([p.detach() for p in ps]) for ps in (self.w_c, self.b_c, self.w_f, self.b_f)]
        d.W_coarse, d.B_coarse, d.W_fine, d.B_fine = [ctypes.cast(a, ctypes.c_void_p) for a in self._arrs]
        if self.mode >= 1:
            d.packed_coarse = r.coarse_model._packed_ptr(self.w_c, self.b_c)
            d.packed_fine = r.fine_model._packed_ptr(self.w_f, self.b_f)
        if self.dt:
            self.C = d.C = int(C)
            d.pixel_intensity_factor = float(r.pixel_intensity_factor)
            self._la = [torch.stack([m.log_absortpion[str(c)].detach().reshape(()) for c in ops.AIA_CHANNELS]).contiguous()
                        for m in (r.coarse_model, r.fine_model)]
            self._vc = [m.volumetric_constant.detach().reshape(1).contiguous() for m in (r.coarse_model, r.fine_model)]
            d.log_abs_coarse, d.log_abs_fine = self._la[0].data_ptr(), self._la[1].data_ptr()
            d.vol_c_coarse, d.vol_c_fine = self._vc[0].data_ptr(), self._vc[1].data_ptr()
            d.table_x, d.table_y = r._table_x.data_ptr(), r._table_y.data_ptr()
        nbytes = _lib.lib().snf_render_ws_bytes(ctypes.byref(d), self.N, int(self.train))
        if nbytes < 0:
            _lib.check(int(nbytes), 'snf_render_ws_bytes')
        if self._ws is None or self._ws.numel() < nbytes + 1024:
            self._ws = torch.empty(nbytes + 1024, device=self.dev, dtype=torch.uint8)
        self.ws_ptr = (self._ws.data_ptr() + 1023) // 1024 * 1024

    def forward(self, rays_o, rays_d, times, wavelengths=None, t_rand=None, reg_grad_scale: float = 0.0) -> Dict[str, torch.Tensor]:
        f32 = lambda x: x.to(self.dev, torch.float32).contiguous()
        rays_o, rays_d, times = f32(rays_o), f32(rays_d), f32(times)
        N = rays_o.shape[0]
        if N != self.N:
            raise SnfError(f'FusedRender was sized for {self.N} rays, got {N}')
        if self.dt and wavelengths is None:
            raise SnfError('the density-temperature head needs wavelengths[N,C]')
        wl = f32(wavelengths) if wavelengths is not None and self.dt else None
        self._refresh(wl.shape[1] if wl is not None else 1)
        if self.perturb and t_rand is None:
            t_rand = torch.rand((N, self.desc.S), device=self.dev)     # the reference's draw (sampling.py:97)
        if not self.perturb:
            t_rand = None
        t_rand = f32(t_rand) if t_rand is not None else None
        S, n_new, C = self.desc.S, self.desc.n_new, self.C
        new = lambda *shape: torch.empty(*shape, device=self.dev, dtype=torch.float32)
        out = {'z_vals_stratified': new(N, S), 'coarse_image': new(N, C), 'z_vals_hierarchical': new(N, n_new),
               'fine_image': new(N, C), 'height_map': new(N), 'absorption_map': new(N), 'regularization': new(N, S + n_new)}
        p = lambda t: t.data_ptr() if t is not None else None
        _lib.check(_lib.lib().snf_render_fused_fwd(
            ctypes.byref(self.desc), p(rays_o), p(rays_d), p(times), p(wl), p(t_rand), N, self.ws_ptr, int(self.train),
            float(reg_grad_scale), p(out['z_vals_stratified']), p(out['coarse_image']), p(out['z_vals_hierarchical']),
            p(out['fine_image']), p(out['height_map']), p(out['absorption_map']), p(out['regularization']), _stream()),
            'snf_render_fused_fwd')
        out['image'] = out['fine_image']
        self._saved = (rays_d, wl)
        return out

    def backward(self, g_coarse_image, g_fine_image, with_reg_grad: bool = True):
        """Gradients of every trainable tensor, as new tensors: {'coarse_model': {'W': [...], 'B': [...], 'log_abs',
        'vol_c'}, 'fine_model': ...}."""
        if not self.train:
            raise SnfError('FusedRender(train=False) kept nothing for a backward')
        rays_d, wl = self._saved
        f32 = lambda x: x.to(self.dev, torch.float32).contiguous()
        g_c, g_f = f32(g_coarse_image), f32(g_fine_image)
        res = {}
        arrs = []
        for name, ws, bs in (('coarse_model', self.w_c, self.b_c), ('fine_model', self.w_f, self.b_f)):
            res[name] = {'W': [torch.empty_like(w) for w in ws], 'B': [torch.empty_like(b) for b in bs],
                         'log_abs': torch.zeros(7, device=self.dev), 'vol_c': torch.zeros(1, device=self.dev)}
            arrs += [_ptr_array(res[name]['W']), _ptr_array(res[name]['B'])]
        p = lambda t: t.data_ptr() if t is not None else None
        _lib.check(_lib.lib().snf_render_fused_bwd(
            ctypes.byref(self.desc), p(rays_d), p(wl), self.N, self.ws_ptr, p(g_c), p(g_f), int(with_reg_grad),
            arrs[0], arrs[1], arrs[2], arrs[3], p(res['coarse_model']['log_abs']), p(res['coarse_model']['vol_c']),
            p(res['fine_model']['log_abs']), p(res['fine_model']['vol_c']), _stream()), 'snf_render_fused_bwd')
        return res
