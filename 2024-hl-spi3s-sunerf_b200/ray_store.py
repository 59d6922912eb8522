"""GPU-resident ray-batch feed.

Replaces the reference's `MmapDataset` + `DataLoader(batch_size=None, shuffle=True, num_workers=...)`
(sunerf/data/dataset.py:7-29, sunerf/data/loader/base_loader.py:41-55): there every training batch is an
`np.copy` of a slice of four memory-mapped `.npy` files made by a worker process, pinned and copied to the GPU, and
Lightning 'dp' scatters it over the GPUs.  Once a step takes a few milliseconds that path is the bottleneck.

Here the same on-disk files (`rays_batches.npy [M,2,3]`, `times_batches.npy [M,1]`, `images_batches.npy [M,C]`,
optionally `wavelengths_batches.npy [M,C]`; single_channel.py:57-72, multi_thermal_loader.py:66-83) are uploaded ONCE
into HBM (28 + 8 C bytes per ray: 100 M rays of a 193 A data set are 3.6 GB of the 180 GB), a batch is a set of
zero-copy views, the epoch order is the permutation of batch indices the reference's DataLoader would draw from the
same torch RNG state, and each rank takes its contiguous share of every global batch (what 'dp' scatter does).
"""
from __future__ import annotations

import os
from typing import Dict, Iterator, List, Optional

import numpy as np
import torch

from . import parallel
from ._lib import SnfError

FILES = {'rays': 'rays_batches.npy', 'time': 'times_batches.npy', 'target_image': 'images_batches.npy',
         'wavelengths': 'wavelengths_batches.npy'}


def dataloader_batch_order(n_batches: int) -> List[int]:
    """The batch order `DataLoader(ds, batch_size=None, shuffle=True)` yields from the current torch RNG state:
    the iterator first draws its base seed, then RandomSampler seeds a private generator for one randperm."""
    torch.empty((), dtype=torch.int64).random_()                      # _BaseDataLoaderIter: base seed (discarded here)
    seed = int(torch.empty((), dtype=torch.int64).random_().item())   # RandomSampler.__iter__
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n_batches, generator=g).tolist()


class RayStore:
    """All training rays of a data set resident in HBM.

        store = RayStore.from_directory(working_dir, batch_size=1024 * world, device='cuda', rank=rank, world=world)
        for idx in store.epoch_order():
            b = store.batch(idx)           # views: rays_o, rays_d [n,3], time [n,1], target_image [n,C] (, wavelengths)
            trainer.step(b['rays_o'], b['rays_d'], b['time'], b['target_image'], b.get('wavelengths'))
    """

    def __init__(self, arrays: Dict[str, np.ndarray], batch_size: int, device, rank: int = 0, world: int = 1):
        if 'rays' not in arrays or 'time' not in arrays or 'target_image' not in arrays:
            raise SnfError("RayStore needs 'rays', 'time' and 'target_image' arrays")
        m = arrays['rays'].shape[0]
        if arrays['rays'].shape[1:] != (2, 3):
            raise SnfError(f"rays must be [M,2,3], got {arrays['rays'].shape}")
        for k, v in arrays.items():
            if v.shape[0] != m:
                raise SnfError(f'{k}: {v.shape[0]} rows, rays has {m}')
        if batch_size % world != 0:
            raise SnfError(f'global batch {batch_size} is not divisible by the world size {world}')
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise SnfError('RayStore keeps the rays in GPU memory: it needs a CUDA device')
        self.n_rays, self.batch_size, self.rank, self.world = m, int(batch_size), rank, world
        self.data: Dict[str, torch.Tensor] = {}
        for k, v in arrays.items():          # one upload per array, float32 on the device (the reference casts per batch)
            self.data[k] = torch.from_numpy(np.array(v, dtype=np.float32, order='C')).to(self.device)   # np.array: one host copy out of the mmap
        rays = self.data.pop('rays')
        self.data['rays_o'] = rays[:, 0].contiguous()      # dataset rays are [origin, direction] (single_channel.py:44)
        self.data['rays_d'] = rays[:, 1].contiguous()
        for k in ('time', 'target_image', 'wavelengths'):
            if k in self.data and self.data[k].dim() == 1:
                self.data[k] = self.data[k][:, None]

    @classmethod
    def from_directory(cls, working_dir: str, batch_size: int, device, rank: int = 0, world: int = 1, mmap: bool = True):
        arrays = {}
        for k, f in FILES.items():
            p = os.path.join(working_dir, f)
            if os.path.exists(p):
                arrays[k] = np.load(p, mmap_mode='r' if mmap else None)
        return cls(arrays, batch_size, device, rank, world)

    def __len__(self) -> int:
        return int(np.ceil(self.n_rays / self.batch_size))           # dataset.py:18-21

    def nbytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.data.values())

    def epoch_order(self, shuffle: bool = True) -> List[int]:
        return dataloader_batch_order(len(self)) if shuffle else list(range(len(self)))

    def shard_range(self, idx: int, rank: Optional[int] = None):
        """Row range [a, b) of global batch `idx` that `rank` owns.  The batch is rows [idx B, (idx+1) B) of the files
        (dataset.py:23-27; the last one may be short) and is split as evenly as `parallel.shard_rows` does: shards differ
        by at most one ray and none is empty while the batch has at least `world` rays.  (torch's 'dp' scatter hands out
        ceil-sized chunks and simply uses fewer replicas on a short tail; one process per GPU cannot drop a rank out of
        the all-reduce, and an empty shard is a no-op step there - see RayTrainer.step.)"""
        if not 0 <= idx < len(self):
            raise IndexError(idx)
        lo, hi = idx * self.batch_size, min((idx + 1) * self.batch_size, self.n_rays)
        sl = parallel.shard_rows(hi - lo, self.rank if rank is None else rank, self.world)
        return lo + sl.start, lo + sl.stop

    def batch(self, idx: int) -> Dict[str, torch.Tensor]:
        """This rank's contiguous share of global batch `idx` as zero-copy views."""
        a, b = self.shard_range(idx)
        return {k: v[a:b] for k, v in self.data.items()}

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        for idx in self.epoch_order():
            yield self.batch(idx)
