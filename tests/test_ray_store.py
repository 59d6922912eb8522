"""N2: the GPU-resident ray-batch feed against the reference's MmapDataset + DataLoader
semantics (sunerf/data/dataset.py:7-29, sunerf/data/loader/base_loader.py:41-55), restated with torch's own DataLoader."""
import os

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader, Dataset


class _IdxDataset(Dataset):          # batch_size=None: the loader yields dataset[idx] for sampler indices
    def __init__(self, n): self.n = n
    def __len__(self): return self.n
    def __getitem__(self, i): return i


@pytest.mark.parametrize('n', [1, 7, 64])
def test_epoch_order_is_what_the_reference_dataloader_draws(n):
    from sunerf_b200.ray_store import dataloader_batch_order
    for seed in (0, 1234):
        torch.manual_seed(seed)
        ref = [int(i) for i in DataLoader(_IdxDataset(n), batch_size=None, shuffle=True, num_workers=0)]
        ref2 = [int(i) for i in DataLoader(_IdxDataset(n), batch_size=None, shuffle=True, num_workers=0)]   # 2nd epoch
        torch.manual_seed(seed)
        assert dataloader_batch_order(n) == ref and dataloader_batch_order(n) == ref2


def _write(tmp_path, m, c=1, wl=False):
    rng = np.random.default_rng(0)
    arrs = {'rays_batches.npy': rng.normal(size=(m, 2, 3)).astype(np.float32),
            'times_batches.npy': rng.uniform(0, 30, (m, 1)),                       # float64 on disk, as np.ones_like(images)*t gives
            'images_batches.npy': rng.uniform(0, 1, (m, c)).astype(np.float32)}
    if wl:
        arrs['wavelengths_batches.npy'] = np.tile(np.array([94., 171., 193.][:c], dtype=np.float32), (m, 1))
    for k, v in arrs.items():
        np.save(os.path.join(tmp_path, k), v)
    return arrs


def test_tail_batch_gives_every_rank_a_shard():
    """A short last batch (here 9 rays over 8 ranks, then 3 over 8) is split evenly: no rank is empty while the batch has
    at least `world` rays (ceil-sized chunks left trailing ranks with nothing and hung the all-reduce), the shards tile
    the batch in rank order, and a batch smaller than the world leaves the surplus ranks an EMPTY shard (a no-op step)."""
    from sunerf_b200.ray_store import RayStore
    for m, B, world in ((1033, 1024, 8), (1027, 1024, 8), (2048 + 49, 1024, 8), (100, 64, 3)):
        st = RayStore.__new__(RayStore)                      # the split needs no device
        st.n_rays, st.batch_size, st.world, st.rank = m, B, world, 0
        for idx in range(len(st)):
            lo, hi = idx * B, min((idx + 1) * B, m)
            rng = [st.shard_range(idx, r) for r in range(world)]
            assert rng[0][0] == lo and rng[-1][1] == hi
            assert all(rng[r][1] == rng[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in rng]
            assert max(sizes) - min(sizes) <= 1
            assert min(sizes) >= 1 or hi - lo < world


@pytest.mark.gpu
def test_ray_store_batches_equal_mmap_dataset_slices(tmp_path):
    import sunerf_b200 as s
    m, B, world = 1000, 256, 2
    arrs = _write(str(tmp_path), m, c=3, wl=True)
    stores = [s.RayStore.from_directory(str(tmp_path), B, 'cuda', rank=r, world=world) for r in range(world)]
    assert len(stores[0]) == 4 and stores[0].nbytes() == m * (24 + 4 + 12 + 12)
    for idx in range(4):
        lo, hi = idx * B, min((idx + 1) * B, m)                                   # dataset.py:23-27
        parts = [st.batch(idx) for st in stores]
        for key, f, sel in (('rays_o', 'rays_batches.npy', 0), ('rays_d', 'rays_batches.npy', 1)):
            got = torch.cat([p[key] for p in parts]).cpu().numpy()
            assert np.array_equal(got, arrs[f][lo:hi, sel])
        for key, f in (('time', 'times_batches.npy'), ('target_image', 'images_batches.npy'), ('wavelengths', 'wavelengths_batches.npy')):
            got = torch.cat([p[key] for p in parts]).cpu().numpy()
            assert np.array_equal(got, arrs[f][lo:hi].astype(np.float32))
        # contiguous, even shares (at most one ray apart)
        assert parts[0]['rays_o'].shape[0] == -(-(hi - lo) // world) and parts[1]['rays_o'].shape[0] == (hi - lo) // world
    # views, not copies
    assert stores[0].batch(0)['rays_o'].data_ptr() == stores[0].data['rays_o'].data_ptr()
    with pytest.raises(IndexError):
        stores[0].batch(4)


@pytest.mark.gpu
def test_ray_store_feeds_the_trainer(tmp_path):
    import sunerf_b200 as s
    rays = s.rays.synthetic_rays(192, seed=2, H=32, W=32, plate_arcsec=70.0)
    np.save(os.path.join(tmp_path, 'rays_batches.npy'), np.stack([rays['rays_o'].numpy(), rays['rays_d'].numpy()], 1))
    np.save(os.path.join(tmp_path, 'times_batches.npy'), rays['times'].numpy())
    np.save(os.path.join(tmp_path, 'images_batches.npy'), rays['target'].numpy())
    store = s.RayStore.from_directory(str(tmp_path), 64, 'cuda')
    torch.manual_seed(0)
    rend = s.EmissionRadiativeTransfer(Rs_per_ds=1, model_config={'d_filter': 64, 'n_layers': 3}).cuda()
    tr = s.RayTrainer(rend)
    losses = [tr.step(b['rays_o'], b['rays_d'], b['time'], b['target_image'])['losses'][0].item() for b in store]
    assert len(losses) == 3 and all(np.isfinite(losses))
    tr.check_finite()
