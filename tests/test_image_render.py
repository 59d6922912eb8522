"""N1: observer rays generated on the device and the full-image render driver.

CPU: the oracle's pose_spherical / image_rays against the golden vectors produced by the reference's own
pose_spherical (coordinate_transformation.py:36-54) and get_rays (ray_sampling.py:7-36) - oracle/make_golden_rays.py.
GPU: the ray kernel against the oracle (float32 ulp level: the direction cosines are double-precision sin/cos of the
device vs glibc), and ObserverRenderer against the oracle render and against itself under batching / row sharding.
"""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import sunerf_oracle as orc


def _cases():
    g = golden('rays.npz')
    n = len([k for k in g.files if k.endswith('.params')])
    return g, n


def test_oracle_rays_match_reference_golden():
    g, n = _cases()
    assert n >= 3
    for c in range(n):
        H, W, plate, lat, lon, dist = g[f'case{c}.params']
        c2w = orc.pose_spherical(-np.deg2rad(lon), np.deg2rad(lat), dist)
        assert np.array_equal(c2w, g[f'case{c}.c2w'])
        ro, rd = orc.image_rays(int(H), int(W), plate, lat, lon, dist)
        assert np.array_equal(ro, g[f'case{c}.rays_o']) and np.array_equal(rd, g[f'case{c}.rays_d'])


def test_host_pose_matches_reference_golden():
    """the product's own host-side pose (image_render.pose_spherical) - no GPU needed"""
    from sunerf_b200.image_render import pose_spherical
    g, n = _cases()
    for c in range(n):
        _, _, _, lat, lon, dist = g[f'case{c}.params']
        assert np.array_equal(pose_spherical(-np.deg2rad(lon), np.deg2rad(lat), dist), g[f'case{c}.c2w'])
    shifted = pose_spherical(0.3, -0.1, 10.0, shift=(1.0, 2.0, 3.0))
    assert np.allclose(shifted[:3, 3] - pose_spherical(0.3, -0.1, 10.0)[:3, 3], [1.0, 2.0, 3.0])


def _ulp_diff(a, b):
    ai, bi = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, np.int64(-2 ** 31) - ai, ai)
    bi = np.where(bi < 0, np.int64(-2 ** 31) - bi, bi)
    return np.abs(ai - bi)


@pytest.mark.gpu
def test_image_rays_kernel_matches_oracle():
    import sunerf_b200 as s
    g, n = _cases()
    cases = [tuple(g[f'case{c}.params']) for c in range(n)] + [(256, 192, 9.4, 5.0, 123.0, orc.R_OBS), (1024, 1024, 2.4, -3.0, 10.0, orc.R_OBS)]
    for H, W, plate, lat, lon, dist in cases:
        H, W = int(H), int(W)
        c2w = s.image_render.pose_spherical(-np.deg2rad(lon), np.deg2rad(lat), dist)
        ro, rd = s.ops.image_rays(c2w, H, W, plate, 'cuda')
        ro_ref, rd_ref = orc.image_rays(H, W, plate, lat, lon, dist)
        assert np.array_equal(ro.cpu().numpy(), ro_ref)
        d = _ulp_diff(rd.cpu().numpy(), rd_ref)
        assert d.max() <= 1, d.max()                       # device vs glibc double sin/cos, after rounding to float32
        assert (d == 0).mean() >= 0.999, (d == 0).mean()
        # a row block is the same rays (multi-GPU row sharding)
        r0, r1 = H // 3, H // 3 + max(1, H // 4)
        _, rd_blk = s.ops.image_rays(c2w, H, W, plate, 'cuda', first=r0 * W, count=(r1 - r0) * W)
        assert torch.equal(rd_blk, rd[r0 * W:r1 * W])


@pytest.mark.gpu
def test_observer_renderer_against_oracle_and_under_sharding():
    import sunerf_b200 as s
    from conftest import oracle_params
    H, W, plate = 12, 10, 200.0
    lat, lon, time = np.deg2rad(4.0), np.deg2rad(50.0), 3.25
    torch.manual_seed(5)
    rend = s.EmissionRadiativeTransfer(Rs_per_ds=1, sampling_config={'type': 'stratified', 'perturb': False}).cuda()
    r = s.ObserverRenderer(rend, (H, W), plate)
    out = r.render_observer_image(lat, lon, time, batch_size=4096)
    assert out['fine_image'].shape == (H, W, 1) and out['z_vals_stratified'].shape == (H, W, 64)
    assert out['regularization'].shape == (H, W, 192) and out['height_map'].shape == (H, W)
    # the reference path: host rays -> render (perturb off), through the CPU oracle
    ro, rd = orc.image_rays(H, W, plate, 4.0, 50.0, orc.R_OBS)
    pc, pf = oracle_params(rend.coarse_model), oracle_params(rend.fine_model)
    ref = orc.render(orc.RenderConfig(kind='emission'), pc, pf, torch.from_numpy(ro), torch.from_numpy(rd),
                     torch.full((H * W, 1), time), None, None)
    for k in ('coarse_image', 'fine_image'):
        a, b = out[k].reshape(-1), ref[k].detach().numpy().reshape(-1)
        assert (np.abs(a - b) <= 2e-5 * np.abs(b) + 1e-12).all(), (k, np.abs(a - b).max())
    # ragged batches and row sharding over 3 "ranks" give the same image, bit for bit (rays are independent)
    out_b = r.render_observer_image(lat, lon, time, batch_size=37)
    parts = [r.render_sharded(rank, 3, lat, lon, time, batch_size=64) for rank in range(3)]
    stitched = s.image_render.stitch_rows(parts)
    for k in out:
        assert np.array_equal(out[k], out_b[k]), k
        assert np.array_equal(out[k], stitched[k]), k
    # resolution override keeps the field of view (Map.resample semantics)
    out_half = r.render_observer_image(lat, lon, time, resolution=(6, 5))
    assert out_half['fine_image'].shape == (6, 5, 1)


@pytest.mark.gpu
def test_observer_renderer_density_temperature_channels():
    """render_mhd.yaml shape of call: SimpleStar field, 6 AIA channels, wavelengths broadcast over the batch."""
    import sunerf_b200 as s
    H, W, plate = 8, 6, 400.0
    wl = [94, 171, 193, 211, 304, 335]
    rend = s.DensityTemperatureRadiativeTransfer(Rs_per_ds=1, model=s.SimpleStar, pixel_intensity_factor=1e10,
                                                 sampling_config={'type': 'stratified', 'perturb': False}).cuda()
    with torch.no_grad():
        for m in (rend.coarse_model, rend.fine_model):
            for c in orc.AIA_CHANNELS:
                m.log_absortpion[str(c)].fill_(1e-6)      # the default 19-20 makes the star opaque: image == 0
    r = s.ObserverRenderer(rend, (H, W), plate)
    out = r.render_observer_image(0.1, 2.0, 0.0, wl=wl, batch_size=20)
    assert out['fine_image'].shape == (H, W, 6) and np.isfinite(out['fine_image']).all() and (out['fine_image'] > 0).any()
    c2w = s.image_render.pose_spherical(-2.0, 0.1, s.rays.R_OBS)
    ro, rd = s.ops.image_rays(c2w, H, W, plate, 'cuda')
    with torch.no_grad():
        ref = rend(ro, rd, torch.zeros(H * W, 1, device='cuda'), torch.tensor(wl, dtype=torch.float32, device='cuda').repeat(H * W, 1))
    assert np.array_equal(out['fine_image'].reshape(-1, 6), ref['fine_image'].cpu().numpy())
