"""Stream-ordered Python entry points over the C ABI: one function per exported kernel group.

Every function takes/returns torch CUDA tensors (PyTorch owns all memory), launches on the current torch CUDA
stream and never synchronises.  No function here has a CPU implementation: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

AIA_CHANNELS = (94, 131, 171, 193, 211, 304, 335)
# field-network modes of the C ABI: fp32 SIMT (any shape), 16-bit tensor cores, split-precision tensor cores (8 x 512 only)
MLP_MODES = {'fp32': 0, 'bf16': 1, 'x3': 2}
TC_MODES = ('bf16', 'x3')


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.SnfError(f'{name}: sunerf_b200 kernels need CUDA tensors (no CPU fallback exists)')
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _ptr_array(ts: Sequence[torch.Tensor]):
    return (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


# ------------------------------------------------------------------------------------------ sampling
def stratified_sample(rays_o, rays_d, t_vals, t_rand, distance: float, solar_R: float, want_points: bool = False,
                      spherical: bool = False):
    """a1 - StratifiedSampler.forward (sunerf/train/sampling.py:68-102), or SphericalSampler.forward (:16-54) with
    spherical=True.  Returns (z_vals[N,S], points or None)."""
    rays_o, rays_d = _f32(rays_o, 'rays_o'), _f32(rays_d, 'rays_d')
    t_vals = _f32(t_vals, 't_vals').reshape(-1)
    N, S = rays_o.shape[0], t_vals.numel()
    if t_rand is not None:
        t_rand = _f32(t_rand, 't_rand')
        assert tuple(t_rand.shape) == (N, S)
    z = torch.empty(N, S, device=rays_o.device, dtype=torch.float32)
    pts = torch.empty(N, S, 3, device=rays_o.device, dtype=torch.float32) if want_points else None
    fn = _lib.lib().snf_spherical_sample if spherical else _lib.lib().snf_stratified_sample
    _lib.check(fn(rays_o.data_ptr(), rays_d.data_ptr(), t_vals.data_ptr(), _ptr(t_rand), N, S, float(distance), float(solar_R),
                  z.data_ptr(), _ptr(pts), _stream()), 'snf_spherical_sample' if spherical else 'snf_stratified_sample')
    return z, pts


def hier_resample(z_vals, weights, u, cdf_in=None, want_inds: bool = False, want_cdf: bool = False, per_ray_u: bool = False):
    """a2 - HierarchicalSampler (sampling.py:111-169). Returns (new_z[N,n], z_comb[N,S+n], inds|None, cdf|None).
    u: the shared ascending grid linspace(0,1,n) (perturb=False), or with per_ray_u=True the [N,n] torch.rand draws of
    perturb=True (sampling.py:144-146)."""
    z_vals, u = _f32(z_vals, 'z_vals'), _f32(u, 'u')
    N, S = z_vals.shape
    n_new = u.shape[-1] if per_ray_u else u.numel()
    if per_ray_u and tuple(u.shape) != (N, n_new):
        raise _lib.SnfError(f'per-ray u must be [N, n_new] = [{N}, {n_new}], got {tuple(u.shape)}')
    weights = _f32(weights, 'weights') if weights is not None else None
    cdf_in = _f32(cdf_in, 'cdf_in') if cdf_in is not None else None
    dev = z_vals.device
    new_z = torch.empty(N, n_new, device=dev, dtype=torch.float32)
    z_comb = torch.empty(N, S + n_new, device=dev, dtype=torch.float32)
    inds = torch.empty(N, n_new, device=dev, dtype=torch.int64) if want_inds else None
    cdf = torch.empty(N, S - 1, device=dev, dtype=torch.float32) if want_cdf else None
    fn = _lib.lib().snf_hier_resample_perturb if per_ray_u else _lib.lib().snf_hier_resample
    _lib.check(fn(z_vals.data_ptr(), _ptr(weights), u.data_ptr(), _ptr(cdf_in), N, S, n_new, new_z.data_ptr(), z_comb.data_ptr(),
                  _ptr(inds), _ptr(cdf), _stream()), 'snf_hier_resample')
    return new_z, z_comb, inds, cdf


def image_rays(c2w, H: int, W: int, plate_arcsec: float, device, first: int = 0, count: Optional[int] = None,
               center_pixel: Optional[Tuple[float, float]] = None):
    """N1 - get_rays (sunerf/data/ray_sampling.py:7-36) on the device for the regular grid Tx=(j-cx)p, Ty=(i-cy)p.
    c2w: the 4x4 (or 3x4) pose_spherical matrix on the HOST. Returns rays_o, rays_d [count,3] for pixels
    [first, first+count) of the row-major H x W image."""
    import numpy as np
    m = np.ascontiguousarray(np.asarray(c2w, dtype=np.float32))
    if m.shape not in ((4, 4), (3, 4)):
        raise _lib.SnfError('image_rays: c2w must be 4x4 or 3x4')
    count = H * W - first if count is None else count
    cx, cy = ((W - 1) / 2, (H - 1) / 2) if center_pixel is None else center_pixel
    dev = torch.device(device)
    if dev.type != 'cuda':
        raise _lib.SnfError('image_rays: sunerf_b200 kernels need a CUDA device (no CPU fallback exists)')
    with torch.cuda.device(dev):
        rays_o = torch.empty(count, 3, device=dev, dtype=torch.float32)
        rays_d = torch.empty(count, 3, device=dev, dtype=torch.float32)
        _lib.check(_lib.lib().snf_image_rays(m.ctypes.data, H, W, float(plate_arcsec), float(np.pi / 180 / 3600), float(cx),
                                             float(cy), int(first), int(count), rays_o.data_ptr(), rays_d.data_ptr(),
                                             _stream()), 'snf_image_rays')
    return rays_o, rays_d


def make_query(rays_o, rays_d, z, times):
    """a3 - query[N,S,4] = (o + d*z, t) (sampling.py:100, base_tracing.py:64-65)."""
    rays_o, rays_d, z, times = _f32(rays_o, 'rays_o'), _f32(rays_d, 'rays_d'), _f32(z, 'z'), _f32(times, 'times')
    N, S = z.shape
    q = torch.empty(N, S, 4, device=z.device, dtype=torch.float32)
    _lib.check(_lib.lib().snf_make_query(rays_o.data_ptr(), rays_d.data_ptr(), z.data_ptr(), times.data_ptr(), N, S,
                                         q.data_ptr(), _stream()), 'snf_make_query')
    return q


# ------------------------------------------------------------------------------------------ field network
class MLPWorkspace:
    """Per-call scratch of the field network (activations kept for the backward when train=True)."""

    def __init__(self, M: int, n_hidden: int, d_filter: int, mode: str, train: bool, device):
        self.M, self.n_hidden, self.d_filter, self.mode, self.train = M, n_hidden, d_filter, mode, train
        nbytes = _lib.lib().snf_mlp_ws_bytes(M, n_hidden, d_filter, MLP_MODES[mode], int(train))
        if nbytes < 0:
            _lib.check(int(nbytes), 'snf_mlp_ws_bytes')
        # 1024-byte aligned base (UMMA/TMA images)
        self._raw = torch.empty(nbytes + 1024, device=device, dtype=torch.uint8)
        self.ptr = (self._raw.data_ptr() + 1023) // 1024 * 1024


def pack_bytes() -> int:
    return int(_lib.lib().snf_mlp_pack_bytes())


def alloc_packed(device) -> Tuple[torch.Tensor, int]:
    raw = torch.empty(pack_bytes() + 1024, device=device, dtype=torch.uint8)
    return raw, (raw.data_ptr() + 1023) // 1024 * 1024


def mlp_pack_bf16(weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], packed_ptr: int) -> None:
    ws = [_f32(w, 'W') for w in weights]
    bs = [_f32(b, 'b') for b in biases]
    if len(ws) != 9 or tuple(ws[0].shape) != (512, 84) or tuple(ws[1].shape) != (512, 512):
        raise _lib.SnfError('bf16 tensor-core path is built for 8 hidden layers of width 512 (the reference default)')
    _lib.check(_lib.lib().snf_mlp_pack_bf16(_ptr_array(ws), _ptr_array(bs), packed_ptr, _stream()), 'snf_mlp_pack_bf16')


def mlp_forward(x, weights, biases, out_offsets=(0.0, 0.0), mode: str = 'fp32', train: bool = False, packed_ptr=None):
    """a4-a6 - encoding + sine MLP (model.py:44-57, 123-132, 169-187).  x[M,4] -> (out[M,2], workspace)."""
    x = _f32(x, 'x')
    M = x.shape[0]
    n_hidden, d = len(weights) - 1, weights[0].shape[0]
    out = torch.empty(M, 2, device=x.device, dtype=torch.float32)
    ws = MLPWorkspace(M, n_hidden, d, mode, train, x.device)
    L = _lib.lib()
    if M == 0:
        return out, ws
    if mode == 'fp32':
        wl = [_f32(w, 'W') for w in weights]
        bl = [_f32(b, 'b') for b in biases]
        _lib.check(L.snf_mlp_fwd_f32(x.data_ptr(), M, _ptr_array(wl), _ptr_array(bl), n_hidden, d, float(out_offsets[0]),
                                     float(out_offsets[1]), out.data_ptr(), ws.ptr, int(train), _stream()), 'snf_mlp_fwd_f32')
    elif mode in TC_MODES:
        if packed_ptr is None:
            raise _lib.SnfError('the tensor-core modes need packed weights (mlp_pack_bf16)')
        fn = L.snf_mlp_fwd_bf16 if mode == 'bf16' else L.snf_mlp_fwd_x3
        _lib.check(fn(x.data_ptr(), M, packed_ptr, float(out_offsets[0]), float(out_offsets[1]), out.data_ptr(), ws.ptr,
                      int(train), _stream()), 'snf_mlp_fwd_' + mode)
    else:
        raise ValueError(f'unknown MLP mode {mode}')
    return out, ws


def mlp_backward(x, weights, grad_out, ws: MLPWorkspace, grad_weights, grad_biases, packed_ptr=None) -> None:
    """Backward of mlp_forward(train=True): OVERWRITES grad_weights / grad_biases (lists of tensors)."""
    x, grad_out = _f32(x, 'x'), _f32(grad_out, 'grad_out')
    L = _lib.lib()
    if ws.mode == 'fp32':
        wl = [_f32(w, 'W') for w in weights]
        _lib.check(L.snf_mlp_bwd_f32(x.data_ptr(), ws.M, _ptr_array(wl), ws.n_hidden, ws.d_filter, grad_out.data_ptr(),
                                     ws.ptr, _ptr_array(grad_weights), _ptr_array(grad_biases), _stream()), 'snf_mlp_bwd_f32')
    else:
        fn = L.snf_mlp_bwd_bf16 if ws.mode == 'bf16' else L.snf_mlp_bwd_x3
        _lib.check(fn(x.data_ptr(), ws.M, packed_ptr, grad_out.data_ptr(), ws.ptr, _ptr_array(grad_weights),
                      _ptr_array(grad_biases), _stream()), 'snf_mlp_bwd_' + ws.mode)


def simple_star(x, rho_0: float, h0: float, T0: float, R_s: float, t_photosphere: float):
    """a7 - SimpleStar.forward (sunerf/model/stellar_model.py:53-102)."""
    x = _f32(x, 'x')
    out = torch.empty(x.shape[0], 2, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().snf_simple_star_fwd(x.data_ptr(), x.shape[0], rho_0, h0, T0, R_s, t_photosphere, out.data_ptr(),
                                              _stream()), 'snf_simple_star_fwd')
    return out


# ------------------------------------------------------------------------------------------ compositing
def composite_emission_fwd(raw, z, rays_d):
    """a8 - emission.py:14-54.  Returns image[N,1], weights[N,S], absorption[N,S]."""
    raw, z, rays_d = _f32(raw, 'raw'), _f32(z, 'z'), _f32(rays_d, 'rays_d')
    N, S = z.shape
    image = torch.empty(N, 1, device=z.device, dtype=torch.float32)
    weights = torch.empty(N, S, device=z.device, dtype=torch.float32)
    absorption = torch.empty(N, S, device=z.device, dtype=torch.float32)
    _lib.check(_lib.lib().snf_composite_emission_fwd(raw.data_ptr(), z.data_ptr(), rays_d.data_ptr(), N, S, image.data_ptr(),
                                                     weights.data_ptr(), absorption.data_ptr(), _stream()),
               'snf_composite_emission_fwd')
    return image, weights, absorption


def composite_emission_bwd(raw, z, rays_d, g_image, g_absorption=None):
    raw, z, rays_d, g_image = _f32(raw, 'raw'), _f32(z, 'z'), _f32(rays_d, 'rays_d'), _f32(g_image, 'g_image')
    g_absorption = _f32(g_absorption, 'g_absorption') if g_absorption is not None else None
    N, S = z.shape
    g_raw = torch.empty(N, S, 2, device=z.device, dtype=torch.float32)
    _lib.check(_lib.lib().snf_composite_emission_bwd(raw.data_ptr(), z.data_ptr(), rays_d.data_ptr(), N, S, g_image.data_ptr(),
                                                     _ptr(g_absorption), g_raw.data_ptr(), _stream()),
               'snf_composite_emission_bwd')
    return g_raw


def composite_dt_fwd(inferences, z, wavelengths, log_abs, vol_c, table_x, table_y, F: float):
    """a9 - density_temperature.py:192-271.  Returns image[N,C], weights[N,S], regq[N,S]."""
    inferences, z, wavelengths = _f32(inferences, 'inferences'), _f32(z, 'z'), _f32(wavelengths, 'wavelengths')
    log_abs, vol_c = _f32(log_abs, 'log_abs').reshape(-1), _f32(vol_c, 'vol_c').reshape(-1)
    N, S = z.shape
    C = wavelengths.shape[1]
    dev = z.device
    image = torch.empty(N, C, device=dev, dtype=torch.float32)
    weights = torch.empty(N, S, device=dev, dtype=torch.float32)
    regq = torch.empty(N, S, device=dev, dtype=torch.float32)
    _lib.check(_lib.lib().snf_composite_dt_fwd(inferences.data_ptr(), z.data_ptr(), wavelengths.data_ptr(), N, S, C,
                                               log_abs.data_ptr(), vol_c.data_ptr(), table_x.data_ptr(), table_y.data_ptr(),
                                               float(F), image.data_ptr(), weights.data_ptr(), regq.data_ptr(), _stream()),
               'snf_composite_dt_fwd')
    return image, weights, regq


def composite_dt_bwd(inferences, z, wavelengths, log_abs, vol_c, table_x, table_y, F: float, g_image, g_regq=None,
                     g_log_abs=None, g_vol_c=None):
    """Returns g_inferences[N,S,2], g_log_abs[7], g_vol_c[1] (the last two accumulate into the given tensors)."""
    inferences, z, wavelengths = _f32(inferences, 'inferences'), _f32(z, 'z'), _f32(wavelengths, 'wavelengths')
    log_abs, vol_c = _f32(log_abs, 'log_abs').reshape(-1), _f32(vol_c, 'vol_c').reshape(-1)
    g_image = _f32(g_image, 'g_image')
    g_regq = _f32(g_regq, 'g_regq') if g_regq is not None else None
    N, S = z.shape
    C = wavelengths.shape[1]
    dev = z.device
    g_inf = torch.empty(N, S, 2, device=dev, dtype=torch.float32)
    if g_log_abs is None:
        g_log_abs = torch.zeros(7, device=dev, dtype=torch.float32)
    if g_vol_c is None:
        g_vol_c = torch.zeros(1, device=dev, dtype=torch.float32)
    _lib.check(_lib.lib().snf_composite_dt_bwd(inferences.data_ptr(), z.data_ptr(), wavelengths.data_ptr(), N, S, C,
                                               log_abs.data_ptr(), vol_c.data_ptr(), table_x.data_ptr(), table_y.data_ptr(),
                                               float(F), g_image.data_ptr(), _ptr(g_regq), g_inf.data_ptr(),
                                               g_log_abs.data_ptr(), g_vol_c.data_ptr(), _stream()), 'snf_composite_dt_bwd')
    return g_inf, g_log_abs, g_vol_c


def render_epilogue(rays_o, rays_d, z_comb, weights, q, r0: float, kind: int, grad_scale: float = 0.0, want_gq: bool = False):
    """a10 - base_tracing.py:91-111 (+ regularization).  Returns height_map[N], absorption_map[N], reg[N,S], g_q|None."""
    rays_o, rays_d, z_comb = _f32(rays_o, 'rays_o'), _f32(rays_d, 'rays_d'), _f32(z_comb, 'z')
    weights, q = _f32(weights, 'weights'), _f32(q, 'q')
    N, S = z_comb.shape
    dev = z_comb.device
    hm = torch.empty(N, device=dev, dtype=torch.float32)
    am = torch.empty(N, device=dev, dtype=torch.float32)
    reg = torch.empty(N, S, device=dev, dtype=torch.float32)
    gq = torch.empty(N, S, device=dev, dtype=torch.float32) if want_gq else None
    _lib.check(_lib.lib().snf_render_epilogue(rays_o.data_ptr(), rays_d.data_ptr(), z_comb.data_ptr(), weights.data_ptr(),
                                              q.data_ptr(), N, S, float(r0), int(kind), hm.data_ptr(), am.data_ptr(),
                                              reg.data_ptr(), float(grad_scale), _ptr(gq), _stream()), 'snf_render_epilogue')
    return hm, am, reg, gq


def train_loss(coarse, fine, target, reg, asinh_scaling: bool, asinh_a: float = 0.005, lambda_image: float = 1.0,
               lambda_reg: float = 1.0, finite_flag: Optional[torch.Tensor] = None):
    """a11 - sunerf.py:98-131 / :173-206.  Returns losses[4]={total,coarse,fine,reg}, g_coarse, g_fine, finite_flag."""
    coarse, fine, target, reg = _f32(coarse, 'coarse'), _f32(fine, 'fine'), _f32(target, 'target'), _f32(reg, 'reg')
    N = coarse.shape[0]
    C = coarse.numel() // N
    dev = coarse.device
    losses = torch.empty(4, device=dev, dtype=torch.float32)
    gc, gf = torch.empty_like(coarse), torch.empty_like(fine)
    if finite_flag is None:
        finite_flag = torch.zeros(1, device=dev, dtype=torch.int32)
    _lib.check(_lib.lib().snf_train_loss(coarse.data_ptr(), fine.data_ptr(), target.data_ptr(), reg.data_ptr(), N, C,
                                         reg.numel(), int(asinh_scaling), float(asinh_a), float(lambda_image),
                                         float(lambda_reg), losses.data_ptr(), gc.data_ptr(), gf.data_ptr(),
                                         finite_flag.data_ptr(), _stream()), 'snf_train_loss')
    return losses, gc, gf, finite_flag


def adam_step(params, grads, exp_avg, exp_avg_sq, step: int, lr: float, scratch, norm_out, beta1=0.9, beta2=0.999, eps=1e-8,
              clip_norm: float = 0.5, grad_scale: float = 1.0) -> None:
    """a12 - clip-by-global-norm + Adam on one flat fp32 buffer (sunerf.py:30-35, run_emission.py:72)."""
    _lib.check(_lib.lib().snf_adam_step(params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                        params.numel(), float(lr), beta1, beta2, eps, int(step), float(clip_norm),
                                        float(grad_scale), scratch.data_ptr(), norm_out.data_ptr(), _stream()), 'snf_adam_step')


def adam_step_sched(params, grads, exp_avg, exp_avg_sq, sched, scratch, norm_out, beta1=0.9, beta2=0.999, eps=1e-8,
                    clip_norm: float = 0.5, grad_scale: float = 1.0) -> None:
    """adam_step with {lr, step, gamma, lr_floor} (float64[4]) on the device: replayable inside a CUDA graph."""
    _lib.check(_lib.lib().snf_adam_step_sched(params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                              params.numel(), sched.data_ptr(), beta1, beta2, eps, float(clip_norm),
                                              float(grad_scale), scratch.data_ptr(), norm_out.data_ptr(), _stream()),
               'snf_adam_step_sched')


def launch_count() -> int:
    return int(_lib.lib().snf_launch_count())
