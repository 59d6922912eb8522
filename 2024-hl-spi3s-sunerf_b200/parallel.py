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


class ReplicatedRendering(torch.nn.Module):
    """Single-process rendering over several GPUs with the call surface of `nn.DataParallel(rendering)` (the reference's
    render path, sunerf/evaluation/loader.py:37-39,143-144): rays are split along dim 0, every device renders its share
    with ITS OWN resident copy of the module, the outputs are concatenated on the first device.  No collective and no
    per-call parameter broadcast - `nn.DataParallel` re-replicates the module on every forward through an NCCL broadcast,
    which deadlocks on torch 2.11 / NCCL 2.28 as soon as two threads enter it at once, exactly what the reference's
    thread-pool render loop (loader.py:226-229) does.  Safe to call from several threads; `sync_weights()` after the
    parameters of `rendering` changed (training, load_state_dict)."""

    def __init__(self, rendering: torch.nn.Module, device_ids: Optional[List[int]] = None):
        super().__init__()
        import copy
        if device_ids is None:
            device_ids = list(range(torch.cuda.device_count()))
        src = next(rendering.parameters()).device
        if src.type != 'cuda' or src.index != device_ids[0]:
            raise ValueError(f'the rendering module must live on the first device (cuda:{device_ids[0]}), it is on {src}')
        self.device_ids = list(device_ids)
        self.module = rendering
        self.replicas = torch.nn.ModuleList([copy.deepcopy(rendering).to(torch.device('cuda', d)) for d in device_ids[1:]])

    @torch.no_grad()
    def sync_weights(self) -> None:
        sd = self.module.state_dict()
        for rep in self.replicas:
            rep.load_state_dict(sd)          # copy_ in place: the replicas' packed weight images notice the new version

    def forward(self, *inputs, **kwargs):
        import threading
        n = inputs[0].shape[0]
        mods = [self.module] + list(self.replicas)
        world = min(len(mods), max(n, 1))
        if world == 1:
            return self.module(*inputs, **kwargs)
        dev0 = torch.device('cuda', self.device_ids[0])
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev0))           # the inputs as the caller's stream produced them
        grad = torch.is_grad_enabled()
        results, errors = [None] * world, [None] * world

        def work(i):
            dev = torch.device('cuda', self.device_ids[i])
            sl = shard_rows(n, i, world)
            try:
                with torch.cuda.device(dev), torch.set_grad_enabled(grad):
                    st = torch.cuda.current_stream(dev)
                    st.wait_event(ready)
                    move = lambda t: t[sl].to(dev, non_blocking=True) if torch.is_tensor(t) and t.dim() > 0 and t.shape[0] == n else t
                    out = mods[i](*[move(t) for t in inputs], **{k: move(v) for k, v in kwargs.items()})
                    out = {k: v.to(dev0, non_blocking=True) for k, v in out.items()}      # stream-ordered peer copies
                    done = torch.cuda.Event()
                    done.record(st)
                    results[i] = (out, done)
            except BaseException as e:  # noqa: BLE001  (re-raised in the caller's thread)
                errors[i] = e

        threads = [threading.Thread(target=work, args=(i,)) for i in range(1, world)]
        for t in threads:
            t.start()
        work(0)
        for t in threads:
            t.join()
        for e in errors:
            if e is not None:
                raise e
        cur = torch.cuda.current_stream(dev0)
        for _, done in results:
            cur.wait_event(done)
        return {k: torch.cat([r[0][k] for r in results], dim=0) for k in results[0][0]}
