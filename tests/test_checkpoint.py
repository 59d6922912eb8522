"""N4: checkpoints written by the reference load into this package.

Fixtures come from the reference itself (oracle/make_golden_ckpt.py): a save_state()-layout `.snf` whose 'rendering' is
the reference's EmissionRadiativeTransfer pickled by class path, a Lightning-layout `.ckpt` state_dict, and the
reference module's outputs for those weights.
"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden, param_digest, rel_err, t


def test_save_state_snf_loads_without_the_reference_package():
    import sunerf_b200 as s
    assert not any(m == 'sunerf' or m.startswith('sunerf.') for m in sys.modules), 'the reference must not be importable here'
    state = s.checkpoint.load_save_state(os.path.join(GOLDEN, 'ref_save_state.snf'))
    r = state['rendering']
    assert isinstance(r, s.EmissionRadiativeTransfer) and isinstance(r.fine_model, s.NeRF)
    assert state['Rs_per_ds'] == 1 and state['seconds_per_dt'] == 86400.0 and state['data_config']['wavelength'] == 193
    assert r.sampler.perturb is False and r.sampler.t_vals.shape == (1, 64) and r.sampler_hierarchical.n_samples == 128
    assert r.fine_model.in_layer[1].weight.shape == (64, 84) and len(r.fine_model.layers) == 3
    assert param_digest(r) == str(golden('ckpt_expected.npz')['digest'])     # every tensor, bit for bit


def test_density_temperature_save_state_loads_without_xitorch():
    """The reference DT module pickles xitorch Interp1D objects (self.response) next to the networks: they become opaque
    stand-ins, the rebuilt module carries NeRF_DT weights, the 7 absorption scalars and the volumetric constants."""
    import sunerf_b200 as s
    assert 'xitorch' not in sys.modules
    state = s.checkpoint.load_save_state(os.path.join(GOLDEN, 'ref_save_state_dt.snf'))
    r = state['rendering']
    assert isinstance(r, s.DensityTemperatureRadiativeTransfer) and isinstance(r.fine_model, s.NeRF_DT)
    assert r.pixel_intensity_factor == 1e17 and r.sampler.perturb is False
    assert abs(float(r.fine_model.log_absortpion['335'].detach()) - 7e-6) < 1e-12
    assert param_digest(r) == str(golden('ckpt_expected_dt.npz')['digest'])


def test_lightning_ckpt_loads_into_a_fresh_module():
    import sunerf_b200 as s
    r = s.EmissionRadiativeTransfer(Rs_per_ds=1, sampling_config={'type': 'stratified', 'perturb': False},
                                    model_config={'d_filter': 64, 'n_layers': 4})
    s.checkpoint.load_lightning_checkpoint(r, os.path.join(GOLDEN, 'ref_lightning.ckpt'))
    assert param_digest(r) == str(golden('ckpt_expected.npz')['digest'])
    with pytest.raises(s.SnfError):
        s.checkpoint.rendering_state_dict({'state_dict': {'foo.bar': torch.zeros(1)}})


def test_save_state_round_trip(tmp_path):
    import sunerf_b200 as s
    torch.manual_seed(3)
    r = s.EmissionRadiativeTransfer(Rs_per_ds=1, model_config={'d_filter': 32, 'n_layers': 3})
    p = str(tmp_path / 'sub' / 'save_state.snf')
    s.checkpoint.save_state(r, p, data_config={'wavelength': 171}, seconds_per_dt=3600.0, ref_time='2012-11-01T00:00:00')
    state = s.checkpoint.load_save_state(p)
    assert param_digest(state['rendering']) == param_digest(r) and state['data_config'] == {'wavelength': 171}


@pytest.mark.gpu
def test_loaded_reference_checkpoint_renders_like_the_reference():
    import sunerf_b200 as s
    g = golden('ckpt_expected.npz')
    r = s.checkpoint.load_save_state(os.path.join(GOLDEN, 'ref_save_state.snf'), precision='fp32')['rendering'].cuda()
    with torch.no_grad():
        out = r(t(g['rays_o']), t(g['rays_d']), t(g['times']))
    assert torch.equal(out['z_vals_stratified'].cpu(), torch.from_numpy(g['out.z_vals_stratified']))
    for k in ('coarse_image', 'fine_image'):
        assert rel_err(out[k], g['out.' + k]) <= 2e-5, (k, rel_err(out[k], g['out.' + k]))
    assert (out['regularization'].cpu() - torch.from_numpy(g['out.regularization'])).abs().max() <= 1e-6


@pytest.mark.gpu
def test_loaded_reference_dt_checkpoint_renders_like_the_reference():
    import sunerf_b200 as s
    g = golden('ckpt_expected_dt.npz')
    r = s.checkpoint.load_save_state(os.path.join(GOLDEN, 'ref_save_state_dt.snf'), precision='fp32')['rendering'].cuda()
    with torch.no_grad():
        out = r(t(g['rays_o']), t(g['rays_d']), t(g['times']), t(g['wavelengths']))
    for k in ('coarse_image', 'fine_image'):
        ref = torch.from_numpy(g['out.' + k])
        assert ((out[k].cpu() - ref).abs() <= 4e-5 * ref.abs() + 1e-12).all(), (k, rel_err(out[k], ref, 1e-9))
    assert (out['fine_image'].cpu()[20:, 0] == 0).all()          # STEREO-masked channels render exactly 0
