"""Pin the oracle's observer-ray synthesis (oracle.pose_spherical / oracle.image_rays) against the reference's own
functions - sunerf/train/coordinate_transformation.py:36-54 (pose_spherical) and sunerf/data/ray_sampling.py:7-36
(get_rays) - executed from /root/reference, and write tests/golden/rays.npz.  Build container only.

The pixel coordinates (Tx, Ty) come from sunpy/astropy WCS code in the reference (all_coordinates_from_map, absent
here); get_rays itself only needs objects with `.Tx/.Ty.to_value(u.rad)`, which this script supplies for the regular
plate-scale grid of SURVEY.md section 8d.
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
import make_golden as mg            # astropy.units shim
from oracle import sunerf_oracle as orc
mg.install_shims()
sys.path.insert(0, '/root/reference')
from sunerf.train.coordinate_transformation import pose_spherical as ref_pose     # noqa: E402
from sunerf.data.ray_sampling import get_rays as ref_get_rays                    # noqa: E402
import astropy.units as u                                                        # the shim


class _Angle:
    def __init__(self, rad): self.rad = rad
    def to_value(self, unit): return self.rad


class _Coords:
    def __init__(self, Tx, Ty): self.Tx, self.Ty = _Angle(Tx), _Angle(Ty)


out = {}
cases = [(12, 10, 40.0, 3.5, 77.0, orc.R_OBS), (8, 8, 600.0, -6.0, 301.25, 180.0), (5, 7, 2.4, 0.0, 0.0, orc.R_OBS)]
for n, (H, W, plate, lat, lon, dist) in enumerate(cases):
    c2w_ref = ref_pose(-np.deg2rad(lon), np.deg2rad(lat), dist).numpy()
    c2w = orc.pose_spherical(-np.deg2rad(lon), np.deg2rad(lat), dist)
    assert np.array_equal(c2w_ref, c2w), (n, np.abs(c2w_ref - c2w).max())
    ii, jj = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
    asec = np.pi / 180 / 3600
    Tx, Ty = (jj - (W - 1) / 2) * plate * asec, (ii - (H - 1) / 2) * plate * asec
    ro_ref, rd_ref = ref_get_rays(_Coords(Tx, Ty), c2w_ref)
    ro, rd = orc.image_rays(H, W, plate, lat, lon, dist)
    assert np.array_equal(ro_ref.reshape(-1, 3).astype(np.float32), ro), n
    assert np.array_equal(rd_ref.reshape(-1, 3).astype(np.float32), rd), n
    out[f'case{n}.params'] = np.array([H, W, plate, lat, lon, dist], dtype=np.float64)
    out[f'case{n}.c2w'] = c2w_ref.astype(np.float32)
    out[f'case{n}.rays_o'] = ro_ref.reshape(-1, 3).astype(np.float32)
    out[f'case{n}.rays_d'] = rd_ref.reshape(-1, 3).astype(np.float32)
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'rays.npz'), **out)
print('oracle == reference (bit for bit) for', len(cases), 'observer poses; wrote tests/golden/rays.npz')
