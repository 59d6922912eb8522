"""Multi-GPU plumbing: one process per GPU, rays sharded by rank, ONE exchange step per training step (an
all-reduce of the flat fp32 gradient buffer), no collective at all for rendering.

Replaces the reference's single-process Lightning 'dp' / nn.DataParallel strategy (sunerf/run_emission.py:64-69,
sunerf/evaluation/loader.py:37-39,143-144), which re-broadcasts all parameters and gathers outputs every step.
Equal shards + sum-all-reduce + 1/world scaling == DP's mean of per-replica mean losses.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Join the torchrun rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*). Returns (rank, world, local_rank)."""
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {'device_id': torch.device('cuda', local)} if backend == 'nccl' else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_slice(n_global: int, rank: int, world: int) -> slice:
    """Rank r owns rays [r*n, (r+1)*n) of every global batch (single_channel.py:67-68 multiplies the batch by N_GPUS)."""
    if n_global % world != 0:
        raise ValueError(f'global ray batch {n_global} is not divisible by world size {world}')
    n = n_global // world
    return slice(rank * n, (rank + 1) * n)


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    n = next(iter(batch.values())).shape[0]
    sl = shard_slice(n, rank, world)
    return {k: v[sl] for k, v in batch.items()}


def shard_rows(n_rows: int, rank: int, world: int) -> slice:
    """Rendering: contiguous row blocks per rank (ragged allowed), no collective; the host stitches the image."""
    base, rem = divmod(n_rows, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def allreduce_sum_async(flat_grad: torch.Tensor, lo: int, hi: int, group=None):
    """Launch the all-reduce of one contiguous gradient bucket; returns a work handle (None when world == 1)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    if os.environ.get('SNF_DEBUG_NO_ALLREDUCE') == '1':      # measurement aid: what the exchange step costs (wrong gradients!)
        return None
    return dist.all_reduce(flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True)


def wait_all(handles: List) -> None:
    for h in handles:
        if h is not None:
            h.wait()
