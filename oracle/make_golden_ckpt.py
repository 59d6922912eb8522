"""Checkpoint-interop fixtures (N4), produced by EXECUTING THE REFERENCE in the build container:
  tests/golden/ref_save_state.snf   - save_state()-layout file (sunerf/model/sunerf.py:62-74) whose 'rendering' is the
                                      reference's own EmissionRadiativeTransfer (small network), pickled by class path
  tests/golden/ref_lightning.ckpt   - {'state_dict': {...}} with the key layout of a Lightning checkpoint of
                                      EmissionSuNeRFModule ('rendering.*' + 'image_scaling.*', sunerf.py:16-27, 87-96)
  tests/golden/ckpt_expected.npz    - rays + the reference module's outputs (with the two emission adapters of
                                      SURVEY.md section 0.1) for those weights
"""
import os, sys
import numpy as np
import torch
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
import make_golden as mg
mg.install_shims()
sys.path.insert(0, mg.REF)
os.chdir(mg.REF)
from sunerf.model.model import NeRF                                   # noqa: E402
from sunerf.rendering.base_tracing import SuNeRFRendering             # noqa: E402
from sunerf.rendering.emission import EmissionRadiativeTransfer       # noqa: E402
from sunerf.train.scaling import ImageAsinhScaling                     # noqa: E402

gold = os.path.join(ROOT, 'tests', 'golden')
torch.manual_seed(21)
rend = EmissionRadiativeTransfer(Rs_per_ds=1, sampling_config={'type': 'stratified', 'perturb': False, 'n_samples': 64},
                                 model_config={'d_filter': 64, 'n_layers': 4})
with torch.no_grad():                                                  # "trained" weights: move away from the init
    for p in rend.parameters():
        p.add_(0.05 * torch.randn_like(p))
# files first (pickled by the reference's class paths, before any adapter touches the classes)
torch.save({'rendering': rend, 'data_config': {'wavelength': 193, 'resolution': [64, 64]}, 'Rs_per_ds': 1,
            'seconds_per_dt': 86400.0, 'ref_time': '2012-08-30T00:00:00'}, os.path.join(gold, 'ref_save_state.snf'))
sd = {'rendering.' + k: v.clone() for k, v in rend.state_dict().items()}
sd.update({'image_scaling.' + k: v.clone() for k, v in ImageAsinhScaling().state_dict().items()})
torch.save({'state_dict': sd, 'epoch': 3, 'global_step': 1234}, os.path.join(gold, 'ref_lightning.ckpt'))

# expected outputs through the reference's own forward (+ the two emission adapters)
_nerf_fwd = NeRF.forward
NeRF.forward = lambda self, x: _nerf_fwd(self, x)['inferences']
SuNeRFRendering.regularization = lambda self, distance, q: torch.relu(distance - 1.2 / self.Rs_per_ds) * (1 - q)
rays = mg.make_rays(48, seed=4)
with torch.no_grad():
    out = rend(rays['rays_o'], rays['rays_d'], rays['times'])
np.savez_compressed(os.path.join(gold, 'ckpt_expected.npz'), rays_o=rays['rays_o'].numpy(), rays_d=rays['rays_d'].numpy(),
                    times=rays['times'].numpy(), digest=np.array(mg.param_digest(rend)),
                    **{'out.' + k: v.numpy() for k, v in out.items()})
print('wrote ref_save_state.snf, ref_lightning.ckpt, ckpt_expected.npz;', {k: tuple(v.shape) for k, v in out.items()})
for f in ('ref_save_state.snf', 'ref_lightning.ckpt', 'ckpt_expected.npz'):
    print(f, os.path.getsize(os.path.join(gold, f)), 'bytes')


# ---------------------------------------------------------------------------------------- density-temperature module
# The reference DT module keeps xitorch Interp1D objects in self.response: give the shim class a picklable identity
# under the real module path, as a file written with xitorch installed would have.
import types as _types
from oracle import sunerf_oracle as orc                                # noqa: E402


class Interp1D:                                                        # picklable stand-in for xitorch.interpolate.Interp1D
    def __init__(self, x, y, method='linear', extrap=0):
        assert method == 'linear' and extrap == 0
        self.x, self.y = x, y

    def __call__(self, xq):
        return orc.interp1d_linear(self.x, self.y, xq)


Interp1D.__module__, Interp1D.__qualname__ = 'xitorch.interpolate', 'Interp1D'
sys.modules['xitorch.interpolate'].Interp1D = Interp1D
import sunerf.rendering.density_temperature as _dtmod                  # noqa: E402
_dtmod.Interp1D = Interp1D
from sunerf.model.model import NeRF_DT                                  # noqa: E402

NeRF.forward = _nerf_fwd                                                # NeRF_DT.forward expects the dict form (model.py:169-187)
torch.manual_seed(22)
rend_dt = _dtmod.DensityTemperatureRadiativeTransfer(
    Rs_per_ds=1, sampling_config={'type': 'stratified', 'perturb': False}, model=NeRF_DT,
    model_config={'d_filter': 64, 'n_layers': 4}, device=torch.device('cpu'), pixel_intensity_factor=1e17)
with torch.no_grad():
    for m_ in (rend_dt.coarse_model, rend_dt.fine_model):
        for i, c in enumerate(orc.AIA_CHANNELS):
            m_.log_absortpion[str(c)].fill_(1e-6 * (1 + i))
torch.save({'rendering': rend_dt, 'data_config': {'wavelengths': list(orc.AIA_CHANNELS)}, 'Rs_per_ds': 1,
            'seconds_per_dt': 86400.0, 'ref_time': '2012-11-01T00:00:00'}, os.path.join(gold, 'ref_save_state_dt.snf'))
rays = mg.make_rays(40, seed=6)
wl = torch.tensor([94., 131., 171., 193., 211., 304., 335.]).repeat(40, 1)
wl[20:] = torch.tensor([0., 0., 171., 193., 211., 304., 0.])
with torch.no_grad():
    out = rend_dt(rays['rays_o'], rays['rays_d'], rays['times'], wl)
np.savez_compressed(os.path.join(gold, 'ckpt_expected_dt.npz'), rays_o=rays['rays_o'].numpy(), rays_d=rays['rays_d'].numpy(),
                    times=rays['times'].numpy(), wavelengths=wl.numpy(), digest=np.array(mg.param_digest(rend_dt)),
                    **{'out.' + k: v.numpy() for k, v in out.items()})
print('wrote ref_save_state_dt.snf', os.path.getsize(os.path.join(gold, 'ref_save_state_dt.snf')), 'bytes;',
      'fine_image range', float(out['fine_image'].min()), float(out['fine_image'].max()))
