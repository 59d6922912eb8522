"""CPU suite: the oracle restatement against the golden vectors produced by the imported reference
(oracle/make_golden.py).  Bit-exact: the oracle issues the same ATen ops as the reference."""
import numpy as np
import torch

from conftest import golden
from oracle import sunerf_oracle as orc


def _eq(a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(np.asarray(b))
    assert a.shape == b.shape
    assert bool(((a == b) | (a.isnan() & b.isnan())).all()), (a.double() - b.double()).abs().max()


def test_stratified_sampler_bit_exact():
    g = golden('sampling.npz')
    o, d = torch.from_numpy(g['rays_o']), torch.from_numpy(g['rays_d'])
    out = orc.stratified_sample(o, d, torch.from_numpy(g['t_vals']), torch.from_numpy(g['t_rand']),
                                torch.tensor(g['distance']), torch.tensor(g['solar_R']))
    _eq(out['z_vals'], g['z_vals'])
    _eq(out['points'], g['points'])
    out = orc.stratified_sample(o, d, torch.from_numpy(g['t_vals']), None, torch.tensor(g['distance']), torch.tensor(g['solar_R']))
    _eq(out['z_vals'], g['z_vals_noperturb'])
    # both hit and miss rays are present
    far = g['z_vals_noperturb'][:, -1]
    assert (far < 215.5).any() and (far > 216.0).any()


def test_hierarchical_resampler_bit_exact():
    g = golden('sampling.npz')
    h = orc.hier_resample(torch.from_numpy(g['rays_o']), torch.from_numpy(g['rays_d']), torch.from_numpy(g['z_vals']),
                          torch.from_numpy(g['weights']))
    _eq(h['inds'], g['inds'])
    _eq(h['new_z_samples'], g['new_z'])
    _eq(h['z_vals'], g['z_comb'])
    _eq(h['points'], g['points_fine'])
    assert g['inds'].min() >= 0 and g['inds'].max() <= 63
    # stage boundary (cdf,u)->inds with the stored cdf
    h2 = orc.hier_resample(torch.from_numpy(g['rays_o']), torch.from_numpy(g['rays_d']), torch.from_numpy(g['z_vals']),
                           torch.from_numpy(g['weights']), cdf_override=torch.from_numpy(g['cdf']))
    _eq(h2['inds'], g['inds'])


def test_field_networks_bit_exact():
    g = golden('field.npz')
    torch.manual_seed(int(g['seed']))
    import sunerf_b200
    net, net_dt = sunerf_b200.NeRF(), sunerf_b200.NeRF_DT()
    from conftest import param_digest, oracle_params
    # same construction order + same seed => same weights and same state_dict keys as the reference modules
    assert param_digest(net) == str(g['digest'])
    assert param_digest(net_dt) == str(g['digest_dt'])
    x = torch.from_numpy(g['x'])
    _eq(orc.positional_encoding(x), g['enc'])
    _eq(orc.field_mlp(x, oracle_params(net)), g['y'])
    _eq(orc.field_mlp(x, oracle_params(net_dt, True), 10.0, 5.0), g['y_dt'])
    _eq(orc.simple_star(torch.from_numpy(g['xs'])), g['ys'])


def _emission_setup():
    g = golden('emission_render.npz')
    torch.manual_seed(int(g['seed']))
    import sunerf_b200
    rend = sunerf_b200.EmissionRadiativeTransfer(Rs_per_ds=1)
    return g, rend


def test_emission_render_and_train_step_bit_exact():
    from conftest import param_digest, oracle_params
    g, rend = _emission_setup()
    assert param_digest(rend) == str(g['digest'])      # state_dict keys, buffers and init identical to the reference
    pc, pf = oracle_params(rend.coarse_model).requires_grad_(), oracle_params(rend.fine_model).requires_grad_()
    o, d, tm = (torch.from_numpy(g[k]) for k in ('rays_o', 'rays_d', 'times'))
    out = orc.render(orc.RenderConfig(kind='emission'), pc, pf, o, d, tm, None, torch.from_numpy(g['t_rand']),
                     keep_intermediates=True)
    for k in ('z_vals_stratified', 'coarse_image', 'z_vals_hierarchical', 'fine_image', 'image', 'height_map',
              'absorption_map', 'regularization'):
        _eq(out[k].detach(), g['out.' + k])
    _eq(out['_inds'], g['inds'])
    _eq(out['_raw_coarse'].detach().reshape(-1, 2), g['raw_c'])
    losses = orc.training_loss(out, torch.from_numpy(g['target']), 'emission')
    _eq(losses['loss'].detach(), g['loss'])
    losses['loss'].backward()
    _eq(pc.weights[-1].grad, g['coarse_model.out_layer.weight.gslice'])
    _eq(pf.weights[4].grad[100:108, 200:216], g['fine_model.layers.3.weight.gslice'])


def test_dt_render_and_train_step():
    from conftest import param_digest, oracle_params
    g = golden('dt_render.npz')
    torch.manual_seed(int(g['seed']))
    import sunerf_b200
    rend = sunerf_b200.DensityTemperatureRadiativeTransfer(Rs_per_ds=1, model=sunerf_b200.NeRF_DT, pixel_intensity_factor=1e17)
    assert param_digest(rend) == str(g['digest'])
    pc, pf = oracle_params(rend.coarse_model, True), oracle_params(rend.fine_model, True)
    pc.log_abs, pf.log_abs = torch.from_numpy(g['log_abs_c']).clone(), torch.from_numpy(g['log_abs_f']).clone()
    pc.requires_grad_(); pf.requires_grad_()
    a = golden('aia_response.npz')
    cfg = orc.RenderConfig(kind='dt', pixel_intensity_factor=1e17, table_x=torch.from_numpy(a['logT']),
                           table_y=torch.from_numpy(a['table']))
    o, d, tm, wl = (torch.from_numpy(g[k]) for k in ('rays_o', 'rays_d', 'times', 'wavelengths'))
    out = orc.render(cfg, pc, pf, o, d, tm, wl, torch.from_numpy(g['t_rand']))
    for k in ('coarse_image', 'fine_image', 'regularization', 'height_map'):
        _eq(out[k].detach(), g['out.' + k])
    # absent channels (wavelength 0, the STEREO mask) render exactly 0
    assert (g['out.fine_image'][len(wl) // 2:, [0, 1, 6]] == 0).all()
    losses = orc.training_loss(out, torch.from_numpy(g['target']), 'dt')
    _eq(losses['loss'].detach(), g['loss'])
    losses['loss'].backward()
    assert np.allclose(pf.vol_c.grad.numpy(), g['fine_model.volumetric_constant.g'], rtol=1e-6)
    assert np.allclose(pc.log_abs.grad.numpy()[3], g['coarse_model.log_absortpion.193.g'], rtol=1e-6)


def test_simple_star_render():
    g = golden('simple_star_render.npz')
    a = golden('aia_response.npz')
    la = torch.from_numpy(g['log_abs'])
    sp = orc.FieldParams([], [], la, torch.tensor(1.0))
    cfg = orc.RenderConfig(kind='dt', pixel_intensity_factor=1e10, table_x=torch.from_numpy(a['logT']),
                           table_y=torch.from_numpy(a['table']), field='simple_star')
    out = orc.render(cfg, sp, sp, *(torch.from_numpy(g[k]) for k in ('rays_o', 'rays_d', 'times', 'wavelengths')),
                     torch.from_numpy(g['t_rand']))
    for k in ('coarse_image', 'fine_image', 'z_vals_hierarchical'):
        _eq(out[k], g['out.' + k])


def test_interp1d_semantics():
    x = torch.tensor([4.0, 4.05, 4.1, 4.15])
    y = torch.tensor([1.0, 2.0, 4.0, 8.0])
    q = torch.tensor([3.9, 4.0, 4.025, 4.05, 4.14, 4.15, 4.2], requires_grad=True)
    v = orc.interp1d_linear(x, y, q)
    assert v[0] == 0 and v[-1] == 0                      # extrap = 0 outside the grid
    assert v[1] == 1.0 and abs(v[2].item() - 1.5) < 1e-5 and abs(v[5].item() - 8.0) < 1e-4
    v.sum().backward()
    assert q.grad[0] == 0 and q.grad[-1] == 0            # no gradient outside
    assert abs(q.grad[2].item() - 20.0) < 1e-3           # slope of the active segment


def test_adam_step_matches_torch():
    torch.manual_seed(0)
    p = [torch.randn(10, requires_grad=True), torch.randn(3, 4, requires_grad=True)]
    st = orc.AdamState(p)
    for q in p:
        q.grad = torch.randn_like(q) * 3
    before = [q.detach().clone() for q in p]
    gn = st.step()
    assert gn > 0.5 and all(not torch.equal(a, b) for a, b in zip(before, p))
    # first Adam step moves every coordinate by ~lr regardless of scale
    assert abs((before[0] - p[0].detach()).abs().max().item() - 1e-4) < 1e-6


def test_optional_samplers_match_reference_golden():
    """SphericalSampler (sampling.py:4-54) and HierarchicalSampler(perturb=True) (:144-146): the oracle restatements
    against vectors produced by the reference's own classes (oracle/make_golden_samplers.py)."""
    g = golden('samplers_optional.npz')
    tt = lambda k: torch.from_numpy(g[k])
    same = lambda a, b: bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())
    out = orc.spherical_sample(tt('sph.rays_o'), tt('sph.rays_d'), tt('sph.t_vals'), tt('sph.t_rand'), tt('sph.distance'), tt('sph.solar_R'))
    assert same(out['z_vals'], tt('sph.z_vals')) and same(out['points'], tt('sph.points'))
    out = orc.spherical_sample(tt('sph.rays_o'), tt('sph.rays_d'), tt('sph.t_vals'), None, tt('sph.distance'), tt('sph.solar_R'))
    assert same(out['z_vals'], tt('sph.z_vals_noperturb'))
    h = orc.hier_resample(tt('hp.rays_o'), tt('hp.rays_d'), tt('hp.z_vals'), tt('hp.weights'), 128, u_rand=tt('hp.u'))
    assert same(h['new_z_samples'], tt('hp.new_z')) and same(h['z_vals'], tt('hp.z_comb')) and same(h['cdf'], tt('hp.cdf'))
    assert bool((h['inds'] == tt('hp.inds')).all())
