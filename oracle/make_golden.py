"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE (imported from /root/reference, never
copied) and pin oracle/sunerf_oracle.py against it, bit for bit where the arithmetic is in-tree.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py

Shims installed before importing the reference:
  * astropy.units      - only unit constants/conversions are used (stellar_model.py:8-31,
                         density_temperature.py:231): a 40-line dimensional Quantity.
  * sunpy.io.special.read_genx - byte-layout reader restated in the oracle.
  * xitorch.interpolate.Interp1D - oracle.interp1d_linear (PARITY UNPINNED, see oracle header).
  * pytorch_lightning.LightningModule - nn.Module with a no-op .log, so the reference's own
    training_step (sunerf/model/sunerf.py:98-131, 173-206) can be called directly.
  * sunerf.data.loader.base_loader - dummy (imports sunpy/pandas I/O; not on the path).
Adapters applied to the reference emission path (it raises at HEAD without them):
  * NeRF.forward returns the tensor instead of {'inferences': tensor}
  * SuNeRFRendering.regularization is elementwise [N,S] (no [:, :, None])
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference'
sys.path.insert(0, ROOT)
from oracle import sunerf_oracle as orc  # noqa: E402


# ------------------------------------------------------------------ shims
class _Q:
    """value * m^a * K^b ; enough of astropy.units for the two call sites."""

    def __init__(self, value, m=0, K=0, si=1.0):
        self.value, self.m, self.K, self.si = value, m, K, si   # si: SI value of ONE unit

    def _same(self, o):
        return self.m == o.m and self.K == o.K

    def __rmul__(self, v):
        return _Q(v * self.value, self.m, self.K, self.si)

    __mul__ = lambda self, o: _Q(self.value * o.value, self.m + o.m, self.K + o.K, self.si * o.si) \
        if isinstance(o, _Q) else _Q(self.value * o, self.m, self.K, self.si)

    def __truediv__(self, o):
        return _Q(self.value / o.value, self.m - o.m, self.K - o.K, self.si / o.si)

    def __rtruediv__(self, v):
        return _Q(v / self.value, -self.m, -self.K, 1.0 / self.si)

    def __pow__(self, p):
        return _Q(self.value ** p, self.m * p, self.K * p, self.si ** p)

    def to(self, o):
        assert self._same(o), 'unit mismatch'
        return _Q(self.value * self.si / o.si / o.value, o.m, o.K, o.si)

    def to_value(self, o):
        return self.to(o).value


def install_shims():
    u = types.ModuleType('astropy.units')
    u.m = _Q(1.0, m=1); u.cm = _Q(1.0, m=1, si=1e-2); u.Mm = _Q(1.0, m=1, si=1e6)
    u.solRad = _Q(1.0, m=1, si=orc.SOLRAD_M); u.K = _Q(1.0, K=1); u.rad = _Q(1.0)
    astropy = types.ModuleType('astropy'); astropy.units = u
    sys.modules.update({'astropy': astropy, 'astropy.units': u})

    def read_genx(path):
        x, y = orc.read_aia_response(path, aia_exp_time=1.0)
        blob = open(path, 'rb').read()
        out = {'HEADER': {}}
        for k, ch in enumerate(orc.AIA_CHANNELS):
            tresp = np.frombuffer(blob, dtype='>f8', count=101, offset=1152 + 1252 * k + 436).astype(np.float64)
            out[f'A{ch}'] = {'LOGTE': x.numpy().copy(), 'TRESP': tresp}
        return out
    sunpy = types.ModuleType('sunpy'); sio = types.ModuleType('sunpy.io'); sp = types.ModuleType('sunpy.io.special')
    sp.read_genx = read_genx; sio.special = sp; sunpy.io = sio
    sys.modules.update({'sunpy': sunpy, 'sunpy.io': sio, 'sunpy.io.special': sp})

    class Interp1D:
        def __init__(self, x, y, method='linear', extrap=0):
            assert method == 'linear' and extrap == 0
            self.x, self.y = x, y

        def __call__(self, xq):
            return orc.interp1d_linear(self.x, self.y, xq)
    xi = types.ModuleType('xitorch'); xii = types.ModuleType('xitorch.interpolate')
    xii.Interp1D = Interp1D; xi.interpolate = xii
    sys.modules.update({'xitorch': xi, 'xitorch.interpolate': xii})

    pl = types.ModuleType('pytorch_lightning')

    class LightningModule(torch.nn.Module):
        def log(self, *a, **k):
            pass
    pl.LightningModule = LightningModule
    sys.modules['pytorch_lightning'] = pl
    bl = types.ModuleType('sunerf.data.loader.base_loader'); bl.BaseDataModule = object
    sys.modules['sunerf.data.loader.base_loader'] = bl


def param_digest(module: torch.nn.Module) -> str:
    h = hashlib.sha1()
    for k, v in module.state_dict().items():
        h.update(k.encode()); h.update(v.detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def to_oracle_params(model, dt=False) -> orc.FieldParams:
    ws = [model.in_layer[1].weight] + [l.weight for l in model.layers] + [model.out_layer.weight]
    bs = [model.in_layer[1].bias] + [l.bias for l in model.layers] + [model.out_layer.bias]
    la = vc = None
    if dt:
        la = torch.stack([model.log_absortpion[str(c)] for c in orc.AIA_CHANNELS])
        vc = model.volumetric_constant
    return orc.FieldParams([w.detach().clone() for w in ws], [b.detach().clone() for b in bs],
                           None if la is None else la.detach().clone(),
                           None if vc is None else vc.detach().clone())


GRAD_SLICES = {  # small, fixed windows of per-parameter gradients stored in the fixtures
    'in_layer.1.weight': (slice(0, 8), slice(0, 84)),
    'in_layer.1.bias': (slice(0, 64),),
    'layers.0.weight': (slice(0, 8), slice(0, 16)),
    'layers.3.weight': (slice(100, 108), slice(200, 216)),
    'layers.6.bias': (slice(0, 64),),
    'out_layer.weight': (slice(0, 2), slice(0, 512)),
    'out_layer.bias': (slice(0, 2),),
}


def grads_summary(prefix, model, out):
    for name, p in model.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        out[f'{prefix}.{name}.gnorm'] = g.double().norm().numpy()
        if name in GRAD_SLICES:
            out[f'{prefix}.{name}.gslice'] = g[GRAD_SLICES[name]].numpy().copy()
        if g.numel() <= 8:
            out[f'{prefix}.{name}.g'] = g.numpy().copy()


def check_equal(name, a, b, exact=True):
    a = a.detach() if torch.is_tensor(a) else torch.as_tensor(a)
    b = b.detach() if torch.is_tensor(b) else torch.as_tensor(b)
    if exact:
        same = torch.equal(a, b) or bool(((a == b) | (a.isnan() & b.isnan())).all())
        assert same, f'oracle != reference at {name}: max abs diff {(a.double() - b.double()).abs().max()}'
    else:
        assert torch.allclose(a, b, rtol=1e-6, atol=0), name
    print(f'  pinned {name}')


def make_rays(n, seed):
    r = orc.synthetic_rays(n, seed=seed, H=64, W=64, plate_arcsec=40.0, t_days=30.0)
    return r


def main():
    install_shims()
    sys.path.insert(0, REF)
    os.chdir(REF)   # density_temperature.py:131 opens the genx relative to the CWD
    from sunerf.model.model import NeRF, NeRF_DT
    from sunerf.train import sampling as ref_sampling
    from sunerf.rendering.base_tracing import SuNeRFRendering
    from sunerf.rendering.emission import EmissionRadiativeTransfer
    from sunerf.rendering.density_temperature import DensityTemperatureRadiativeTransfer
    from sunerf.model.stellar_model import SimpleStar
    from sunerf.model import sunerf as ref_pl

    # ---- the two emission adapters
    _nerf_fwd = NeRF.forward
    NeRF.forward = lambda self, x: _nerf_fwd(self, x)['inferences']
    SuNeRFRendering.regularization = lambda self, distance, q: \
        torch.relu(distance - 1.2 / self.Rs_per_ds) * (1 - q)

    gold = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(gold, exist_ok=True)

    # ================= sampler (a1) + hierarchical resampler (a2)
    print('sampler / resampler')
    rays = make_rays(96, seed=1)
    torch.manual_seed(11)
    samp = ref_sampling.StratifiedSampler(Rs_per_ds=1)
    torch.manual_seed(12)
    t_rand = torch.rand(96, 64)
    torch.manual_seed(12)
    ref = samp(rays['rays_o'], rays['rays_d'])
    mine = orc.stratified_sample(rays['rays_o'], rays['rays_d'], samp.t_vals, t_rand, samp.distance, samp.solar_R)
    check_equal('stratified z_vals', mine['z_vals'], ref['z_vals'])
    check_equal('stratified points', mine['points'], ref['points'])
    samp.perturb = False
    ref_np = samp(rays['rays_o'], rays['rays_d'])
    mine_np = orc.stratified_sample(rays['rays_o'], rays['rays_d'], samp.t_vals, None, samp.distance, samp.solar_R)
    check_equal('stratified (no perturb) z_vals', mine_np['z_vals'], ref_np['z_vals'])
    torch.manual_seed(13)
    # weights: mix of peaky, flat and all-zero rows (empty PDF is an edge case of sample_pdf)
    w = torch.rand(96, 64) ** 4
    w[:8] = 0.0
    w[8:16, 20] = 50.0
    w = w / (w.sum(-1, keepdim=True) + 1e-10)
    hs = ref_sampling.HierarchicalSampler()
    href = hs(rays['rays_o'], rays['rays_d'], ref['z_vals'], w)
    hm = orc.hier_resample(rays['rays_o'], rays['rays_d'], ref['z_vals'], w)
    check_equal('hier new_z', hm['new_z_samples'], href['new_z_samples'])
    check_equal('hier z_combined', hm['z_vals'], href['z_vals'])
    check_equal('hier points', hm['points'], href['points'])
    np.savez_compressed(os.path.join(gold, 'sampling.npz'), rays_o=rays['rays_o'].numpy(), rays_d=rays['rays_d'].numpy(),
                        t_vals=samp.t_vals.numpy(), t_rand=t_rand.numpy(), distance=samp.distance.numpy(),
                        solar_R=samp.solar_R.numpy(), z_vals=ref['z_vals'].numpy(), points=ref['points'].numpy(),
                        z_vals_noperturb=ref_np['z_vals'].numpy(), weights=w.numpy(), cdf=hm['cdf'].numpy(),
                        inds=hm['inds'].numpy(), new_z=href['new_z_samples'].numpy(), z_comb=href['z_vals'].numpy(),
                        points_fine=href['points'].numpy())

    # ================= field MLP (a4-a6) + SimpleStar (a7)
    print('field networks')
    torch.manual_seed(21)
    net = NeRF(); net_dt = NeRF_DT()
    x = torch.cat([torch.randn(192, 3) * 1.2, torch.rand(192, 1) * 30], -1)
    x[:64, :3] += torch.tensor([215.0, 0, 0])   # far-from-origin points exercise big encoder arguments
    y = net(x)
    check_equal('NeRF forward', orc.field_mlp(x, to_oracle_params(net)), y)
    y_dt = net_dt(x)['inferences']
    check_equal('NeRF_DT forward', orc.field_mlp(x, to_oracle_params(net_dt, True), 10.0, 5.0), y_dt)
    star = SimpleStar()
    xs = torch.cat([torch.randn(256, 3) * 0.8, torch.zeros(256, 1)], -1)
    ys = star(xs)['inferences']
    check_equal('SimpleStar forward', orc.simple_star(xs), ys)
    np.savez_compressed(os.path.join(gold, 'field.npz'), seed=21, x=x.numpy(), y=y.detach().numpy(), y_dt=y_dt.detach().numpy(),
                        digest=param_digest(net), digest_dt=param_digest(net_dt), xs=xs.numpy(), ys=ys.detach().numpy(),
                        enc=orc.positional_encoding(x).numpy())

    # ================= emission render + training_step (a8, a10, a11)
    print('emission render + train step')
    N = 40
    rays = make_rays(N, seed=2)
    torch.manual_seed(31)
    mod = ref_pl.EmissionSuNeRFModule(Rs_per_ds=1, seconds_per_dt=1, image_scaling_config={},
                                      validation_dataset_mapping={})
    rend = mod.rendering
    digest = param_digest(rend)
    cap = {}
    rend.coarse_model.register_forward_hook(lambda m, i, o: cap.__setitem__('raw_c', o.detach().clone()))
    rend.fine_model.register_forward_hook(lambda m, i, o: cap.__setitem__('raw_f', o.detach().clone()))
    torch.manual_seed(32)
    t_rand = torch.rand(N, 64)
    torch.manual_seed(32)
    batch = {'tracing': {'rays': torch.stack([rays['rays_o'], rays['rays_d']], 1), 'time': rays['times'],
                         'target_image': rays['target']}}
    loss = mod.training_step(batch, 0)
    loss.backward()
    torch.manual_seed(32)
    with torch.no_grad():
        out = rend(rays['rays_o'], rays['rays_d'], rays['times'])
    cfg = orc.RenderConfig(kind='emission')
    pc, pf = to_oracle_params(rend.coarse_model).requires_grad_(), to_oracle_params(rend.fine_model).requires_grad_()
    mine = orc.render(cfg, pc, pf, rays['rays_o'], rays['rays_d'], rays['times'], None, t_rand, keep_intermediates=True)
    for k in out:
        check_equal(f'emission {k}', mine[k], out[k])
    check_equal('emission raw coarse', mine['_raw_coarse'].reshape(-1, 2), cap['raw_c'])
    check_equal('emission raw fine', mine['_raw_fine'].reshape(-1, 2), cap['raw_f'])
    ml = orc.training_loss(mine, rays['target'], 'emission')
    check_equal('emission loss', ml['loss'], loss)
    ml['loss'].backward()
    check_equal('emission grad out_layer.weight (coarse)', pc.weights[-1].grad, rend.coarse_model.out_layer.weight.grad)
    check_equal('emission grad layers.3.weight (fine)', pf.weights[4].grad, rend.fine_model.layers[3].weight.grad)
    fx = {'seed': 31, 'digest': digest, 'rays_o': rays['rays_o'].numpy(), 'rays_d': rays['rays_d'].numpy(),
          'times': rays['times'].numpy(), 'target': rays['target'].numpy(), 't_rand': t_rand.numpy(),
          'loss': loss.detach().numpy(), 'raw_c': cap['raw_c'].numpy(), 'raw_f': cap['raw_f'].numpy(),
          'cdf': mine['_cdf'].detach().numpy(), 'inds': mine['_inds'].numpy(), 'z_comb': mine['_z_combined'].detach().numpy(),
          'weights_c': mine['_weights_coarse'].detach().numpy()}
    fx.update({f'out.{k}': v.numpy() for k, v in out.items()})
    grads_summary('coarse_model', rend.coarse_model, fx)
    grads_summary('fine_model', rend.fine_model, fx)
    np.savez_compressed(os.path.join(gold, 'emission_render.npz'), **fx)

    # ================= density-temperature render + training_step (a9)
    print('density-temperature render + train step')
    N = 32
    rays = make_rays(N, seed=3)
    torch.manual_seed(41)
    mod = ref_pl.DensityTemperatureSuNeRFModule(Rs_per_ds=1, seconds_per_dt=1, image_scaling_config={}, model=NeRF_DT,
                                                validation_dataset_mapping={}, model_config={})
    rend = mod.rendering
    digest = param_digest(rend)
    with torch.no_grad():   # non-trivial absorption so d/dlog_abs is exercised (init 1e-6 gives optical depth ~0)
        for i, c in enumerate(orc.AIA_CHANNELS):
            rend.coarse_model.log_absortpion[str(c)].fill_(2.0e-6 * (i + 1))
            rend.fine_model.log_absortpion[str(c)].fill_(3.0e-6 * (i + 1))
    wl = torch.tensor(orc.AIA_CHANNELS, dtype=torch.float32)[None].repeat(N, 1)
    wl[N // 2:] = torch.tensor([0, 0, 171, 193, 211, 304, 0], dtype=torch.float32)   # multi_thermal_loader.py:162-168
    target = torch.rand(N, 7, generator=torch.Generator().manual_seed(5))
    torch.manual_seed(42)
    t_rand = torch.rand(N, 64)
    torch.manual_seed(42)
    batch = {'tracing': {'rays': torch.stack([rays['rays_o'], rays['rays_d']], 1), 'time': rays['times'],
                         'target_image': target, 'wavelength': wl}}
    loss = mod.training_step(batch, 0)
    loss.backward()
    torch.manual_seed(42)
    with torch.no_grad():
        out = rend(rays['rays_o'], rays['rays_d'], rays['times'], wl)
    tx, ty = orc.read_aia_response('sunerf/data/aia_temp_resp.genx')
    cfg = orc.RenderConfig(kind='dt', pixel_intensity_factor=1e17, table_x=tx, table_y=ty)
    pc = to_oracle_params(rend.coarse_model, True).requires_grad_()
    pf = to_oracle_params(rend.fine_model, True).requires_grad_()
    mine = orc.render(cfg, pc, pf, rays['rays_o'], rays['rays_d'], rays['times'], wl, t_rand, keep_intermediates=True)
    for k in out:
        check_equal(f'dt {k}', mine[k], out[k])
    ml = orc.training_loss(mine, target, 'dt')
    check_equal('dt loss', ml['loss'], loss)
    ml['loss'].backward()
    ref_la = torch.stack([rend.fine_model.log_absortpion[str(c)].grad if rend.fine_model.log_absortpion[str(c)].grad is not None
                          else torch.tensor(0.) for c in orc.AIA_CHANNELS])
    check_equal('dt grad log_abs (fine)', pf.log_abs.grad, ref_la, exact=False)
    check_equal('dt grad vol_c (coarse)', pc.vol_c.grad, rend.coarse_model.volumetric_constant.grad, exact=False)
    check_equal('dt grad out_layer.weight (fine)', pf.weights[-1].grad, rend.fine_model.out_layer.weight.grad, exact=False)
    fx = {'seed': 41, 'digest': digest, 'rays_o': rays['rays_o'].numpy(), 'rays_d': rays['rays_d'].numpy(),
          'times': rays['times'].numpy(), 'target': target.numpy(), 'wavelengths': wl.numpy(), 't_rand': t_rand.numpy(),
          'loss': loss.detach().numpy(),
          # density_temperature.py:172 calls model.forward() directly (no hooks fire); the oracle's raw
          # outputs are pinned through every downstream quantity checked above
          'raw_c': mine['_raw_coarse'].detach().reshape(-1, 2).numpy(), 'raw_f': mine['_raw_fine'].detach().reshape(-1, 2).numpy(),
          'inds': mine['_inds'].numpy(), 'cdf': mine['_cdf'].detach().numpy(),
          'log_abs_c': pc.log_abs.detach().numpy(), 'log_abs_f': pf.log_abs.detach().numpy()}
    fx.update({f'out.{k}': v.numpy() for k, v in out.items()})
    grads_summary('coarse_model', rend.coarse_model, fx)
    grads_summary('fine_model', rend.fine_model, fx)
    np.savez_compressed(os.path.join(gold, 'dt_render.npz'), **fx)

    # ================= AIA response table (input data of a9; derived from the genx, not source code)
    np.savez_compressed(os.path.join(gold, 'aia_response.npz'), logT=tx.numpy(), table=ty.numpy(),
                        channels=np.array(orc.AIA_CHANNELS))
    # DT render with the analytic SimpleStar field (render_mhd.yaml); log_abs scaled to 1e-6 (SURVEY a7 note)
    print('simple-star DT render')
    rend = DensityTemperatureRadiativeTransfer(Rs_per_ds=1, model=SimpleStar, pixel_intensity_factor=1e10)
    with torch.no_grad():
        for i, c in enumerate(orc.AIA_CHANNELS):
            rend.coarse_model.log_absortpion[str(c)].fill_(1.0e-6 * (i + 1))
            rend.fine_model.log_absortpion[str(c)].fill_(1.0e-6 * (i + 1))
    wl6 = torch.tensor([94, 171, 193, 211, 304, 335], dtype=torch.float32)[None].repeat(N, 1)
    torch.manual_seed(52)
    t_rand = torch.rand(N, 64)
    torch.manual_seed(52)
    with torch.no_grad():
        out = rend(rays['rays_o'], rays['rays_d'], rays['times'], wl6)
    la = torch.tensor([1.0e-6 * (i + 1) for i in range(7)])
    sp = orc.FieldParams([], [], la, torch.tensor(1.0))
    cfg = orc.RenderConfig(kind='dt', pixel_intensity_factor=1e10, table_x=tx, table_y=ty, field='simple_star')
    mine = orc.render(cfg, sp, sp, rays['rays_o'], rays['rays_d'], rays['times'], wl6, t_rand)
    for k in out:
        check_equal(f'simple-star {k}', mine[k], out[k])
    fx = {'rays_o': rays['rays_o'].numpy(), 'rays_d': rays['rays_d'].numpy(), 'times': rays['times'].numpy(),
          'wavelengths': wl6.numpy(), 't_rand': t_rand.numpy(), 'log_abs': la.numpy()}
    fx.update({f'out.{k}': v.numpy() for k, v in out.items()})
    np.savez_compressed(os.path.join(gold, 'simple_star_render.npz'), **fx)
    print('golden fixtures written to', gold)


if __name__ == '__main__':
    main()
