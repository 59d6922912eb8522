"""CPU oracle for the SuNeRF ray-render hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product path (`sunerf_b200`) never does, and fails
loudly when the CUDA extension is missing.

What it is: a plain torch-CPU (fp32) restatement, function by function, of the
reference's PyTorch algorithm for the path SURVEY.md section 8 scopes.  The reference
is pure Python/PyTorch, so "restating it on CPU" means issuing the same ATen
ops in the same order; each function cites the reference file:line it follows
(paths relative to the upstream tree).

Pinning status
  * sampling (a1, a2), encoding + field MLP (a4-a6), emission compositing (a8),
    render orchestration (a10), asinh-MSE loss (a11): PINNED - checked
    bit-for-bit against the reference modules imported in the build container
    by `oracle/make_golden.py`, which also wrote `tests/golden/*.npz`.
    The emission path only runs with the two adapters SURVEY.md section 0.1 describes
    (unwrap ['inferences']; elementwise [N,S] regulariser); the golden script
    applies exactly those two adapters to the imported reference.
  * density-temperature head (a9): the reference calls
    xitorch.interpolate.Interp1D (un-pinned dependency, requirements.txt:7,
    source absent) and sunpy.io.special.read_genx / astropy.units (absent).
    `interp1d_linear` restates xitorch's published LinearInterp1D algorithm;
    `read_aia_response` restates the genx byte layout.  The reference's own DT
    module is executed by make_golden.py with these two restatements injected
    as shims, so everything *around* the interpolation is pinned, but the
    interpolation arithmetic itself is PARITY UNPINNED (no reference test or
    vendored source fixes it).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

AIA_CHANNELS = (94, 131, 171, 193, 211, 304, 335)


# --------------------------------------------------------------------------------------
# a1  StratifiedSampler.forward            sunerf/train/sampling.py:68-102 (ctor :58-66)
# --------------------------------------------------------------------------------------
def stratified_sample(rays_o: torch.Tensor, rays_d: torch.Tensor, t_vals: torch.Tensor,
                      t_rand: Optional[torch.Tensor], distance: torch.Tensor,
                      solar_R: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Bins between |o|-D and |o|+D, far end clipped at the solar surface when the ray hits it.

    `t_rand` is the `torch.rand([N,S])` draw of sampling.py:97 passed in explicitly
    (None == perturb=False).  `distance`, `solar_R`, `t_vals` are the sampler buffers
    (sampling.py:62-66).
    """
    o_sq = rays_o.pow(2).sum(-1)
    r_obs = o_sq.pow(0.5)                                           # :74
    qa = rays_d.pow(2).sum(-1)                                      # :77
    qb = (2 * rays_o * rays_d).sum(-1)                              # :78
    qc = o_sq - solar_R ** 2                                        # :80
    hit = (-qb - torch.sqrt(qb.pow(2) - 4 * qa * qc)) / (2 * qa)    # :81  NaN when the ray misses
    near = r_obs - distance                                         # :83
    far = r_obs + distance                                          # :84
    far = torch.where(torch.isnan(hit), far, hit)                   # :87-88
    z = near[:, None] * (1. - t_vals) + far[:, None] * t_vals       # :90
    if t_rand is not None:                                          # :93-98
        mid = .5 * (z[:, 1:] + z[:, :-1])
        hi = torch.cat([mid, z[:, -1:]], dim=1)
        lo = torch.cat([z[:, :1], mid], dim=1)
        z = lo + (hi - lo) * t_rand
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[:, :, None]   # :100
    return {'points': pts, 'z_vals': z}


def spherical_sample(rays_o: torch.Tensor, rays_d: torch.Tensor, t_vals: torch.Tensor,
                     t_rand: Optional[torch.Tensor], distance: torch.Tensor,
                     solar_R: torch.Tensor) -> Dict[str, torch.Tensor]:
    """SphericalSampler.forward, sampling.py:16-54 (selected by sampling_config {'type': 'spherical'},
    base_tracing.py:27-28): bins between the entry and exit of the sphere of radius `distance`
    (NaN for rays that miss it), far end clipped at the solar surface."""
    qa = rays_d.pow(2).sum(-1)                                      # :24
    qb = (2 * rays_o * rays_d).sum(-1)                              # :25
    qc = rays_o.pow(2).sum(-1) - distance ** 2                      # :26
    near = (-qb - torch.sqrt(qb.pow(2) - 4 * qa * qc)) / (2 * qa)   # :27
    far = (-qb + torch.sqrt(qb.pow(2) - 4 * qa * qc)) / (2 * qa)    # :28
    qc = rays_o.pow(2).sum(-1) - solar_R ** 2                       # :32
    hit = (-qb - torch.sqrt(qb.pow(2) - 4 * qa * qc)) / (2 * qa)    # :33
    far = torch.where(torch.isnan(hit), far, hit)                   # :35-36
    z = near[:, None] * (1. - t_vals) + far[:, None] * t_vals       # :41
    if t_rand is not None:                                          # :44-49
        mid = .5 * (z[:, 1:] + z[:, :-1])
        hi = torch.cat([mid, z[:, -1:]], dim=1)
        lo = torch.cat([z[:, :1], mid], dim=1)
        z = lo + (hi - lo) * t_rand
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[:, :, None]   # :51
    return {'points': pts, 'z_vals': z}


# --------------------------------------------------------------------------------------
# a2  HierarchicalSampler.forward + sample_pdf      sunerf/train/sampling.py:111-169
# --------------------------------------------------------------------------------------
def pdf_to_cdf(weights_inner: torch.Tensor, exact_sum: bool = False) -> torch.Tensor:
    """sampling.py:134-138: pdf=(w+1e-5)/sum(w+1e-5); cdf=[0, cumsum(pdf)].

    `torch.sum` is the ONE platform-dependent quantity of the path: on CPU it is a cascade whose grouping follows the
    vector width of the build, so its last bit differs between CPUs - and from any GPU.
    exact_sum=True (tests only: tie analysis) replaces it by the exactly rounded sum, which is what the CUDA
    resampler computes; everything else is unchanged."""
    w = weights_inner + 1e-5
    norm = w.double().sum(-1, keepdim=True).float() if exact_sum else torch.sum(w, -1, keepdim=True)
    pdf = w / norm
    cdf = torch.cumsum(pdf, dim=-1)
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)


def invert_cdf(bins: torch.Tensor, cdf: torch.Tensor, u: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """sampling.py:147-167: the (cdf,u)->inds stage and the interpolation that follows.

    Returns (samples, inds).  `inds` is the int64 result of searchsorted(right=True): the
    quantity BASELINE.json asks to be bit-exact.
    """
    u = u.expand(list(cdf.shape[:-1]) + [u.shape[-1]]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)                   # :149
    lo_i = torch.clamp(inds - 1, min=0)                             # :152
    hi_i = torch.clamp(inds, max=cdf.shape[-1] - 1)                 # :153
    cdf_lo = torch.gather(cdf, -1, lo_i)
    cdf_hi = torch.gather(cdf, -1, hi_i)
    b_lo = torch.gather(bins, -1, lo_i)
    b_hi = torch.gather(bins, -1, hi_i)
    denom = cdf_hi - cdf_lo                                          # :164
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)  # :165
    t = (u - cdf_lo) / denom                                         # :166
    return b_lo + t * (b_hi - b_lo), inds                            # :167


def hier_resample(rays_o, rays_d, z_vals, weights, n_new: int = 128,
                  cdf_override: Optional[torch.Tensor] = None, exact_sum: bool = False,
                  u_rand: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """sampling.py:111-126.  perturb=False (:141-143): u = linspace(0,1,n_new); perturb=True (:144-146):
    `u_rand[N,n_new]` is the torch.rand draw of :145 passed in explicitly."""
    bins = .5 * (z_vals[..., 1:] + z_vals[..., :-1])                 # :118
    cdf = pdf_to_cdf(weights[..., 1:-1], exact_sum) if cdf_override is None else cdf_override
    u = torch.linspace(0., 1., n_new) if u_rand is None else u_rand
    new_z, inds = invert_cdf(bins, cdf, u)
    new_z = new_z.detach()                                           # :120
    z_comb, _ = torch.sort(torch.cat([z_vals, new_z], dim=-1), dim=-1)  # :123
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_comb[:, :, None]  # :124
    return {'points': pts, 'z_vals': z_comb, 'new_z_samples': new_z, 'cdf': cdf, 'inds': inds,
            'bins': bins}


# --------------------------------------------------------------------------------------
# a4  PositionalEncoding.forward                     sunerf/model/model.py:123-132
# --------------------------------------------------------------------------------------
def positional_encoding(x: torch.Tensor, n_freqs: int = 10, scale_factor: float = 2.) -> torch.Tensor:
    """[x, sin(x_c*f/scale) (f major, c minor), cos(same)] -> 4*(1+2*10)=84 columns."""
    f = (2. ** torch.linspace(0., n_freqs - 1, n_freqs))[None, :, None]     # :112
    arg = x[:, None, :] * f / scale_factor
    return torch.cat([x, torch.sin(arg).reshape(x.shape[0], -1),
                      torch.cos(arg).reshape(x.shape[0], -1)], dim=-1)


# --------------------------------------------------------------------------------------
# a5/a6  NeRF.forward / NeRF_DT.forward               sunerf/model/model.py:44-57, 169-187
# --------------------------------------------------------------------------------------
class FieldParams:
    """Weights of one field network in nn.Linear layout (weight[out,in]); names follow the
    reference state_dict: in_layer.1, layers.0..6, out_layer."""

    def __init__(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
                 log_abs: Optional[torch.Tensor] = None, vol_c: Optional[torch.Tensor] = None):
        self.weights = list(weights)
        self.biases = list(biases)
        self.log_abs = log_abs          # [7] in AIA_CHANNELS order (NeRF_DT only)
        self.vol_c = vol_c              # scalar (NeRF_DT only)

    @staticmethod
    def init(seed: int, d_filter: int = 512, n_layers: int = 8, d_in: int = 84, d_out: int = 2,
             dt: bool = False) -> "FieldParams":
        """torch.nn.Linear default init in the reference's construction order (model.py:31-42)."""
        g = torch.Generator().manual_seed(seed)
        dims = [d_in] + [d_filter] * n_layers + [d_out]
        ws, bs = [], []
        for i in range(len(dims) - 1):
            bound = 1.0 / math.sqrt(dims[i])
            ws.append((torch.rand(dims[i + 1], dims[i], generator=g) * 2 - 1) * bound)
            bs.append((torch.rand(dims[i + 1], generator=g) * 2 - 1) * bound)
        la = torch.full((7,), 1.0e-6) if dt else None        # model.py:154-162
        vc = torch.tensor(1.0) if dt else None               # model.py:164
        return FieldParams(ws, bs, la, vc)

    def tensors(self) -> List[torch.Tensor]:
        out = []
        for w, b in zip(self.weights, self.biases):
            out += [w, b]
        if self.log_abs is not None:
            out += [self.log_abs, self.vol_c]
        return out

    def requires_grad_(self, flag: bool = True) -> "FieldParams":
        for t in self.tensors():
            t.requires_grad_(flag)
        return self


def field_mlp(x: torch.Tensor, p: FieldParams, base_log_density: float = 0.0,
              base_log_temperature: float = 0.0) -> torch.Tensor:
    """sin(Linear) x 8, Linear -> [M,2]; NeRF_DT adds the +10/+5 offsets (model.py:180-183)."""
    h = positional_encoding(x)
    n = len(p.weights)
    for i in range(n - 1):
        h = torch.sin(torch.nn.functional.linear(h, p.weights[i], p.biases[i]))  # :50-52 (Sine :71-72, w0=1)
    out = torch.nn.functional.linear(h, p.weights[-1], p.biases[-1])            # :55
    if base_log_density != 0.0 or base_log_temperature != 0.0:
        out = torch.stack([out[:, 0] + base_log_density, out[:, 1] + base_log_temperature], dim=-1)
    return out


# --------------------------------------------------------------------------------------
# a7  SimpleStar.forward                              sunerf/model/stellar_model.py:53-102
# --------------------------------------------------------------------------------------
SOLRAD_M = 6.957e8           # astropy.units.solRad (IAU 2015 nominal)
SIMPLE_STAR = dict(h0=60.0e6 / SOLRAD_M, T0=1.4e6, R_s=1.02, t_photosphere=5777.0, rho_0=3.0e8)
SIMPLE_STAR_LOG_ABS = (20.4, 20.2, 20.0, 19.8, 19.6, 19.4, 19.2)   # stellar_model.py:34-42


def simple_star(x: torch.Tensor, h0=None, T0=None, R_s=None, t_photosphere=None, rho_0=None) -> torch.Tensor:
    """Hydrostatic isothermal-corona star: returns stack(ln rho, log10 T)."""
    c = SIMPLE_STAR
    h0 = torch.tensor(c['h0'] if h0 is None else h0, dtype=torch.float32)
    T0 = torch.tensor(c['T0'] if T0 is None else T0, dtype=torch.float32)
    R_s = torch.tensor(c['R_s'] if R_s is None else R_s, dtype=torch.float32)
    rho_0 = torch.tensor(c['rho_0'] if rho_0 is None else rho_0, dtype=torch.float32)
    t_ph = c['t_photosphere'] if t_photosphere is None else t_photosphere
    r = torch.sqrt(x[:, 0] ** 2 + x[:, 1] ** 2 + x[:, 2] ** 2)                  # :73
    inside = r <= 1.0
    rho = torch.where(inside, rho_0.expand_as(r), rho_0 * torch.exp(1 / h0 * (1 / r - 1)))   # :83-85
    rho = torch.log(rho)
    ramp = (r - 1) * ((T0 - t_ph) / (R_s - 1)) + t_ph                            # :93
    T = torch.where(inside, torch.full_like(r, t_ph), torch.where(r <= R_s, ramp, T0.expand_as(r)))
    return torch.stack((rho, torch.log10(T)), dim=-1)


# --------------------------------------------------------------------------------------
# a8  EmissionRadiativeTransfer.raw2outputs           sunerf/rendering/emission.py:14-54
#     cumprod_exclusive                               sunerf/rendering/base_tracing.py:135-156
# --------------------------------------------------------------------------------------
def exclusive_cumprod(t: torch.Tensor) -> torch.Tensor:
    c = torch.cumprod(t, -1)
    return torch.cat([torch.ones_like(c[..., :1]), c[..., :-1]], dim=-1)


def composite_emission(raw: torch.Tensor, z_vals: torch.Tensor, rays_d: torch.Tensor) -> Dict[str, torch.Tensor]:
    dz = z_vals[..., 1:] - z_vals[..., :-1]                                    # :24
    dz = torch.cat([dz[..., :1], dz], dim=-1)                                  # :25
    dz = dz * torch.norm(rays_d[..., None, :], dim=-1)                         # :29
    emitted = torch.exp(raw[..., 0]) * dz                                      # :34
    transmit = torch.exp(-torch.relu(raw[..., 1]) * dz)                        # :37
    through = exclusive_cumprod(transmit + 1e-10)                              # :43
    emerging = emitted * through                                               # :46
    image = emerging.sum(1)[:, None]                                           # :48
    weights = emerging / (emerging.sum(1)[:, None] + 1e-10)                    # :51-52
    return {'image': image, 'weights': weights, 'regularizing_quantity': transmit}


# --------------------------------------------------------------------------------------
# a9  DensityTemperatureRadiativeTransfer              sunerf/rendering/density_temperature.py
# --------------------------------------------------------------------------------------
def read_aia_response(path: str, aia_exp_time: float = 2.9) -> Tuple[torch.Tensor, torch.Tensor]:
    """Restates what density_temperature.py:131-146 obtains from sunpy's read_genx for
    `sunerf/data/aia_temp_resp.genx` (XDR/big-endian IDL genx, 9908 bytes): per channel k in
    AIA_CHANNELS order a block at 1152+1252*k with LOGTE (101 x f32) at +32 and TRESP
    (101 x f64) at +436.  Returns (logT[101] f32, table[7,101] f32 = float32(TRESP*exp_time))."""
    blob = open(path, 'rb').read()
    assert len(blob) == 9908, 'unexpected genx size'
    xs, ys = [], []
    for k, ch in enumerate(AIA_CHANNELS):
        base = 1152 + 1252 * k
        name = blob[base:base + 4].rstrip(b'\x00').decode()
        assert name == f'A{ch}', name
        xs.append(np.frombuffer(blob, dtype='>f4', count=101, offset=base + 32).astype(np.float32))
        tresp = np.frombuffer(blob, dtype='>f8', count=101, offset=base + 436).astype(np.float64)
        ys.append(torch.from_numpy(tresp * aia_exp_time).float())              # :142-144
    for x in xs[1:]:
        assert np.array_equal(x, xs[0])
    return torch.from_numpy(xs[0].copy()), torch.stack(ys)


def interp1d_linear(x: torch.Tensor, y: torch.Tensor, xq: torch.Tensor) -> torch.Tensor:
    """xitorch.interpolate.Interp1D(x, y, method='linear', extrap=0)(xq)   [PARITY UNPINNED]

    Published algorithm (xitorch LinearInterp1D, numel(xq) > numel(x) branch, always taken
    here): idxr = clamp(searchsorted(x, xq, right=False), 1, n-1); idxl = idxr-1;
    slope_i = (y[i+1]-y[i])/(x[i+1]-x[i]); out = y[idxl] + (xq-x[idxl])*slope[idxl];
    queries outside [x0, x_last] return the `extrap` constant 0.  Indices carry no
    gradient, so d out/d xq = slope of the active segment (0 outside).
    """
    n = x.shape[0]
    idxr = torch.clamp(torch.searchsorted(x, xq.detach(), right=False), 1, n - 1)
    idxl = idxr - 1
    slope = (y[1:] - y[:-1]) / (x[1:] - x[:-1])
    val = y[idxl] + (xq - x[idxl]) * slope[idxl]
    outside = (xq < x[0]) | (xq > x[-1])
    return torch.where(outside, torch.zeros_like(val), val)


def composite_dt(inferences: torch.Tensor, z_vals: torch.Tensor, wavelengths: torch.Tensor,
                 log_abs: torch.Tensor, vol_c: torch.Tensor, table_x: torch.Tensor,
                 table_y: torch.Tensor, pixel_intensity_factor: float) -> Dict[str, torch.Tensor]:
    """density_temperature.py:192-271.  `wavelengths[N,C]` float, 0 == channel absent;
    `log_abs[7]` in AIA_CHANNELS order.  Lines :223-232 of the reference (dists / cm) are
    dead code and are not restated."""
    N, S, _ = inferences.shape
    C = wavelengths.shape[1]
    wl = wavelengths[:, None, :].expand(N, S, C)                               # :219
    rho = torch.exp(torch.relu(inferences[..., 0]))[:, :, None].expand(N, S, C)   # :237-238
    theta = torch.relu(inferences[..., 1])[:, :, None].expand(N, S, C)          # :241-242
    resp = torch.zeros(N, S, C)
    kappa = torch.zeros(N, S, C)
    for k, ch in enumerate(AIA_CHANNELS):                                       # :244-256
        sel = (wl == float(ch))
        r_k = interp1d_linear(table_x, table_y[k], theta.reshape(-1)).reshape(N, S, C)
        resp = torch.where(sel, r_k, resp)
        kappa = torch.where(sel, torch.relu(log_abs[k]).expand(N, S, C), kappa)
    absorb = rho * kappa                                                        # :260
    A = torch.cumulative_trapezoid(absorb, x=z_vals[:, :, None], dim=1)         # :261
    emis = rho.pow(2) * resp                                                    # :263
    term = torch.exp(-A) * emis[:, 0:-1, :]                                     # :264
    image = torch.trapezoid(term, x=z_vals[:, 0:-1, None], dim=1) * vol_c * pixel_intensity_factor  # :265
    q = torch.relu(inferences[..., 0])
    weights = q / (q.sum(1)[:, None] + 1e-10)                                   # :268-269
    return {'image': image, 'weights': weights, 'regularizing_quantity': q}    # :271


# --------------------------------------------------------------------------------------
# a3 + a10  SuNeRFRendering.forward                   sunerf/rendering/base_tracing.py:46-111
# --------------------------------------------------------------------------------------
class RenderConfig:
    def __init__(self, Rs_per_ds: float = 1.0, distance: float = 1.3, n_samples: int = 64,
                 n_hier: int = 128, kind: str = 'emission', pixel_intensity_factor: float = 1e10,
                 table_x: Optional[torch.Tensor] = None, table_y: Optional[torch.Tensor] = None,
                 field: str = 'nerf'):
        self.Rs_per_ds = Rs_per_ds
        self.kind = kind                        # 'emission' | 'dt'
        self.field = field                      # 'nerf' | 'simple_star'
        self.n_hier = n_hier
        self.F = pixel_intensity_factor
        self.table_x, self.table_y = table_x, table_y
        # sampler buffers exactly as sampling.py:62-66 builds them
        self.distance = torch.tensor(distance / Rs_per_ds, dtype=torch.float32)
        self.solar_R = torch.tensor(1 / Rs_per_ds, dtype=torch.float32)
        self.t_vals = torch.linspace(0., 1., n_samples)[None]


def _eval_field(cfg: RenderConfig, p: FieldParams, pts: torch.Tensor, times: torch.Tensor) -> torch.Tensor:
    N, S, _ = pts.shape
    q = torch.cat([pts, times[:, None].repeat(1, S, 1)], -1).view(-1, 4)       # :64-65 / :83-84
    if cfg.field == 'simple_star':
        raw = simple_star(q)
    elif cfg.kind == 'dt':
        raw = field_mlp(q, p, 10.0, 5.0)
    else:
        raw = field_mlp(q, p)
    return raw.reshape(N, S, 2)


def _composite(cfg, p, raw, z, rays_d, wavelengths):
    if cfg.kind == 'dt':
        return composite_dt(raw, z, wavelengths, p.log_abs, p.vol_c, cfg.table_x, cfg.table_y, cfg.F)
    return composite_emission(raw, z, rays_d)


def render(cfg: RenderConfig, coarse: FieldParams, fine: FieldParams, rays_o, rays_d, times,
           wavelengths=None, t_rand=None, keep_intermediates: bool = False, exact_sum: bool = False) -> Dict[str, torch.Tensor]:
    s = stratified_sample(rays_o, rays_d, cfg.t_vals, t_rand, cfg.distance, cfg.solar_R)   # :59
    z = s['z_vals']
    raw_c = _eval_field(cfg, coarse, s['points'], times)
    c_out = _composite(cfg, coarse, raw_c, z, rays_d, wavelengths)                         # :68-71
    h = hier_resample(rays_o, rays_d, z, c_out['weights'], cfg.n_hier, exact_sum=exact_sum)  # :76
    raw_f = _eval_field(cfg, fine, h['points'], times)
    f_out = _composite(cfg, fine, raw_f, h['z_vals'], rays_d, wavelengths)                  # :86-89
    q = f_out['regularizing_quantity']
    absorption_map = (1 - q).sum(-1)                                                        # :99
    dist = h['points'].pow(2).sum(-1).pow(0.5)                                              # :101
    height_map = (f_out['weights'] * dist).sum(-1)                                          # :102
    if cfg.kind == 'dt':
        reg = torch.relu(dist - 1.25 / cfg.Rs_per_ds) * torch.relu(q)   # density_temperature.py:273-274
    else:
        # base_tracing.py:43-44 as evidently intended (elementwise [N,S]); see SURVEY.md section 0.1
        reg = torch.relu(dist - 1.2 / cfg.Rs_per_ds) * (1 - q)
    out = {'z_vals_stratified': z, 'coarse_image': c_out['image'],
           'z_vals_hierarchical': h['new_z_samples'], 'fine_image': f_out['image'],
           'image': f_out['image'], 'height_map': height_map, 'absorption_map': absorption_map,
           'regularization': reg}
    if keep_intermediates:
        out.update({'_raw_coarse': raw_c, '_raw_fine': raw_f, '_weights_coarse': c_out['weights'],
                    '_cdf': h['cdf'], '_inds': h['inds'], '_z_combined': h['z_vals'],
                    '_weights_fine': f_out['weights'], '_q_fine': q})
    return out


# --------------------------------------------------------------------------------------
# a11  training_step losses                             sunerf/model/sunerf.py:98-131, 173-206
#      ImageAsinhScaling                                sunerf/train/scaling.py:17-28
# --------------------------------------------------------------------------------------
def asinh_scale(img: torch.Tensor, a: float = 0.005, vmax: float = 1.0) -> torch.Tensor:
    norm = torch.tensor(np.arcsinh(1 / a), dtype=torch.float32)
    return torch.asinh(img / torch.tensor(vmax) / torch.tensor(a, dtype=torch.float32)) / norm


def training_loss(out: Dict[str, torch.Tensor], target: torch.Tensor, kind: str,
                  lambda_image: float = 1.0, lambda_reg: float = 1.0) -> Dict[str, torch.Tensor]:
    for k, v in out.items():                                                    # sunerf.py:105-107
        if k.startswith('_'):
            continue
        assert not torch.isnan(v).any(), f'{k} contains NaN'
        assert not torch.isinf(v).any(), f'{k} contains Inf'
    mse = torch.nn.functional.mse_loss
    if kind == 'emission':
        tgt = asinh_scale(target)
        lc = mse(asinh_scale(out['coarse_image']), tgt)
        lf = mse(asinh_scale(out['fine_image']), tgt)
    else:
        lc = mse(out['coarse_image'], target)
        lf = mse(out['fine_image'], target)
    lr = out['regularization'].mean()
    return {'loss': lambda_image * (lc + lf) + lambda_reg * lr, 'coarse': lc, 'fine': lf, 'reg': lr}


# --------------------------------------------------------------------------------------
# a12  optimiser step: clip-by-global-norm 0.5 (run_emission.py:72), Adam 1e-4, ExponentialLR
#      (sunerf.py:30-40).  Lightning 1.9.3 is absent; this is the plain-torch equivalent.
# --------------------------------------------------------------------------------------
class AdamState:
    def __init__(self, params: List[torch.Tensor], lr: float = 1e-4, lr_end: float = 1e-5, iters: float = 1e6):
        self.params = params
        self.opt = torch.optim.Adam(params, lr=lr)
        self.sched = torch.optim.lr_scheduler.ExponentialLR(self.opt, gamma=(lr_end / lr) ** (1 / iters))

    def step(self, clip: float = 0.5) -> float:
        gn = torch.nn.utils.clip_grad_norm_(self.params, clip)
        self.opt.step()
        if self.sched.get_last_lr()[0] > 5e-5:                                   # sunerf.py:38-39
            self.sched.step()
        return float(gn)


def train_step(cfg, coarse: FieldParams, fine: FieldParams, opt: Optional[AdamState], rays_o, rays_d,
               times, target, wavelengths=None, t_rand=None) -> Dict[str, torch.Tensor]:
    params = coarse.tensors() + fine.tensors()
    for p_ in params:
        p_.grad = None
    out = render(cfg, coarse, fine, rays_o, rays_d, times, wavelengths, t_rand)
    losses = training_loss(out, target, cfg.kind)
    losses['loss'].backward()
    if opt is not None:
        losses['grad_norm'] = opt.step()
    return losses


# --------------------------------------------------------------------------------------
# Synthetic rays: pose_spherical (train/coordinate_transformation.py:36-54)
# + get_rays (data/ray_sampling.py:7-36), restated with numpy.
# --------------------------------------------------------------------------------------
R_OBS = 1.495978707e11 / SOLRAD_M      # 1 AU in solar radii = 215.032


def pose_spherical(theta: float, phi: float, radius: float) -> np.ndarray:
    t = np.eye(4, dtype=np.float32); t[2, 3] = radius
    rp = np.array([[1, 0, 0, 0], [0, np.cos(phi), -np.sin(phi), 0],
                   [0, np.sin(phi), np.cos(phi), 0], [0, 0, 0, 1]], dtype=np.float32)
    rt = np.array([[np.cos(theta), 0, -np.sin(theta), 0], [0, 1, 0, 0],
                   [np.sin(theta), 0, np.cos(theta), 0], [0, 0, 0, 1]], dtype=np.float32)
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
    return flip @ (rt @ (rp @ t))


def image_rays(H: int, W: int, plate_arcsec: float, lat_deg: float, lon_deg: float,
               r_obs: float = R_OBS) -> Tuple[np.ndarray, np.ndarray]:
    c2w = pose_spherical(-np.deg2rad(lon_deg), np.deg2rad(lat_deg), r_obs)
    ii, jj = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
    asec = np.pi / 180 / 3600
    Tx = (jj - (W - 1) / 2) * plate_arcsec * asec
    Ty = (ii - (H - 1) / 2) * plate_arcsec * asec
    d = np.stack([np.sin(Tx), -np.sin(Ty) * np.cos(Tx), -np.cos(Tx) * np.cos(Ty)], -1).astype(np.float32)
    rays_d = np.sum(d[..., None, :] * c2w[:3, :3], axis=-1).astype(np.float32)
    rays_o = np.broadcast_to(c2w[:3, -1], rays_d.shape).astype(np.float32)
    return rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)


def synthetic_rays(n: int, seed: int = 0, H: int = 256, W: int = 256, plate_arcsec: float = 9.4,
                   t_days: float = 30.0, n_views: int = 3) -> Dict[str, torch.Tensor]:
    rng = np.random.default_rng(seed)
    os_, ds_ = [], []
    for v in range(n_views):
        o, d = image_rays(H, W, plate_arcsec, lat_deg=rng.uniform(-7, 7), lon_deg=rng.uniform(0, 360))
        os_.append(o); ds_.append(d)
    o = np.concatenate(os_); d = np.concatenate(ds_)
    sel = rng.permutation(o.shape[0])[:n]
    return {'rays_o': torch.from_numpy(o[sel].copy()), 'rays_d': torch.from_numpy(d[sel].copy()),
            'times': torch.from_numpy(rng.uniform(0, t_days, (n, 1)).astype(np.float32)),
            'target': torch.from_numpy(rng.uniform(0, 1, (n, 1)).astype(np.float32))}
