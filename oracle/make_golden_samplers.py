"""Generate tests/golden/samplers_optional.npz by EXECUTING THE REFERENCE's optional samplers
(imported from /root/reference, never copied) and pin the oracle restatements against them bit for bit:

  * SphericalSampler.forward                       sunerf/train/sampling.py:4-54   (perturb on and off)
  * HierarchicalSampler(perturb=True).forward      sunerf/train/sampling.py:111-169, u = torch.rand (:144-146)

Run in the build container only (the GPU box has no /root/reference):  python oracle/make_golden_samplers.py
RNG protocol: torch.manual_seed(s) right before the reference call; the same seed and the same torch.rand call
(shape [N,64] for the jitter, [N,128] for u) reproduce the draw that is stored and passed to the kernels explicitly.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
from oracle import sunerf_oracle as orc  # noqa: E402

warnings.filterwarnings('ignore')
from sunerf.train.sampling import HierarchicalSampler, SphericalSampler, StratifiedSampler  # noqa: E402


def same(a, b):
    return bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())


def main():
    out = {}
    # rays on a wide field (100 arcsec / pixel, +-3.3 R_sun): some hit the Sun, some only the 2 R_sun sphere, some miss it (NaN rows)
    b = orc.synthetic_rays(96, seed=4, H=64, W=64, plate_arcsec=100.0)
    ro, rd = b['rays_o'], b['rays_d']
    sph = SphericalSampler(Rs_per_ds=1)
    torch.manual_seed(31)
    ref = sph(ro, rd)
    torch.manual_seed(31)
    t_rand = torch.rand(96, 64)
    mine = orc.spherical_sample(ro, rd, sph.t_vals, t_rand, sph.distance, sph.solar_R)
    assert same(mine['z_vals'], ref['z_vals']) and same(mine['points'], ref['points']), 'oracle spherical_sample != reference'
    sph.perturb = False
    ref_np = sph(ro, rd)
    assert same(orc.spherical_sample(ro, rd, sph.t_vals, None, sph.distance, sph.solar_R)['z_vals'], ref_np['z_vals'])
    n_nan = int(torch.isnan(ref['z_vals']).any(-1).sum())
    assert 0 < n_nan < 96, n_nan
    print(f'SphericalSampler: oracle == reference bit for bit ({n_nan} of 96 rays miss the sphere -> NaN rows)')
    out.update({'sph.rays_o': ro.numpy(), 'sph.rays_d': rd.numpy(), 'sph.t_rand': t_rand.numpy(), 'sph.t_vals': sph.t_vals.numpy(),
                'sph.distance': sph.distance.numpy(), 'sph.solar_R': sph.solar_R.numpy(), 'sph.z_vals': ref['z_vals'].numpy(),
                'sph.points': ref['points'].numpy(), 'sph.z_vals_noperturb': ref_np['z_vals'].numpy()})

    # HierarchicalSampler(perturb=True) on stratified depths + peaked weights (narrow pdf bins exercise the denom floor)
    b = orc.synthetic_rays(80, seed=6, H=64, W=64, plate_arcsec=30.0)
    ro, rd = b['rays_o'], b['rays_d']
    strat = StratifiedSampler(Rs_per_ds=1)
    torch.manual_seed(32)
    z = strat(ro, rd)['z_vals']
    g = torch.Generator().manual_seed(33)
    w = torch.rand(80, 64, generator=g) ** 6
    w[::5] = 0.                                       # all-zero weights: uniform pdf
    w = w / (w.sum(-1, keepdim=True) + 1e-10)
    hs = HierarchicalSampler(perturb=True)
    torch.manual_seed(34)
    ref = hs(ro, rd, z, w)
    torch.manual_seed(34)
    u = torch.rand(80, 128)
    mine = orc.hier_resample(ro, rd, z, w, 128, u_rand=u)
    assert same(mine['new_z_samples'], ref['new_z_samples']) and same(mine['z_vals'], ref['z_vals']) and \
        same(mine['points'], ref['points']), 'oracle hier_resample(u_rand) != reference'
    print('HierarchicalSampler(perturb=True): oracle == reference bit for bit')
    out.update({'hp.rays_o': ro.numpy(), 'hp.rays_d': rd.numpy(), 'hp.z_vals': z.numpy(), 'hp.weights': w.numpy(), 'hp.u': u.numpy(),
                'hp.cdf': mine['cdf'].numpy(), 'hp.inds': mine['inds'].numpy(), 'hp.new_z': ref['new_z_samples'].numpy(),
                'hp.z_comb': ref['z_vals'].numpy()})
    path = os.path.join(ROOT, 'tests', 'golden', 'samplers_optional.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
