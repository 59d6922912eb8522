"""CPU suite, world_size 2 over gloo: the N>1 host logic - ray sharding by rank plus ONE sum-all-reduce of the flat
gradient scaled by 1/world - reproduces the full-batch gradient (what Lightning 'dp' computes as the mean of
replica losses, sunerf/run_emission.py:64-69).  Per-rank work is the CPU oracle; the collective is the real one."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    os.environ.update({'RANK': str(rank), 'WORLD_SIZE': str(world), 'LOCAL_RANK': str(rank),
                       'MASTER_ADDR': '127.0.0.1', 'MASTER_PORT': str(port)})
    torch.set_num_threads(2)
    from oracle import sunerf_oracle as orc
    from sunerf_b200 import parallel
    r, w, _ = parallel.init_distributed('gloo')
    assert (r, w) == (rank, world)
    g = golden('emission_render.npz')
    batch = {k: torch.from_numpy(g[k]) for k in ('rays_o', 'rays_d', 'times', 'target', 't_rand')}
    batch = {k: v[:32] for k, v in batch.items()}                     # 32 rays -> 16 per rank
    mine = parallel.shard_batch(batch, rank, world)
    assert mine['rays_o'].shape[0] == 16
    # small networks keep the CPU test fast; same seed on every rank == replicated weights without a broadcast
    pc = orc.FieldParams.init(1, d_filter=32, n_layers=3).requires_grad_()
    pf = orc.FieldParams.init(2, d_filter=32, n_layers=3).requires_grad_()
    cfg = orc.RenderConfig(kind='emission')
    orc.train_step(cfg, pc, pf, None, mine['rays_o'], mine['rays_d'], mine['times'], mine['target'], None, mine['t_rand'])
    flat = torch.cat([t.grad.reshape(-1) for t in pc.tensors() + pf.tensors()])
    n_fine = sum(t.numel() for t in pc.tensors())
    handles = [parallel.allreduce_sum_async(flat, 0, n_fine), parallel.allreduce_sum_async(flat, n_fine, flat.numel())]
    parallel.wait_all(handles)
    flat /= world
    if rank == 0:
        np.save(os.path.join(out_dir, 'dp_grad.npy'), flat.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_ray_shard_allreduce_equals_full_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = torch.from_numpy(np.load(tmp_path / 'dp_grad.npy'))
    from oracle import sunerf_oracle as orc
    g = golden('emission_render.npz')
    b = {k: torch.from_numpy(g[k])[:32] for k in ('rays_o', 'rays_d', 'times', 'target', 't_rand')}
    pc = orc.FieldParams.init(1, d_filter=32, n_layers=3).requires_grad_()
    pf = orc.FieldParams.init(2, d_filter=32, n_layers=3).requires_grad_()
    orc.train_step(orc.RenderConfig(kind='emission'), pc, pf, None, b['rays_o'], b['rays_d'], b['times'], b['target'], None, b['t_rand'])
    ref = torch.cat([t.grad.reshape(-1) for t in pc.tensors() + pf.tensors()])
    # mean over equal shards of per-shard means == full-batch mean (MSE and reg.mean() are both per-ray means)
    assert torch.allclose(got, ref, rtol=2e-4, atol=1e-9), (got - ref).abs().max()


def test_shard_helpers():
    from sunerf_b200 import parallel
    assert parallel.shard_slice(8192, 3, 8) == slice(3072, 4096)
    rows = [parallel.shard_rows(1024 + 3, r, 8) for r in range(8)]
    assert rows[0].start == 0 and rows[-1].stop == 1027
    assert all(a.stop == b.start for a, b in zip(rows, rows[1:]))
    import pytest
    with pytest.raises(ValueError):
        parallel.shard_slice(1000, 0, 3)
