"""N > 1 on real GPUs (-m gpu, skipped with fewer than 2 devices): one process per GPU over NCCL, rays sharded by rank,
one all-reduce of the flat gradient per step.  Replaces Lightning 'dp' (run_emission.py:64-69):
a world-2 step on the two halves of a batch must equal the single-process step on the whole batch, replicas must stay
bit-identical to each other (the wgrad's atomics make each rank's own gradient non-deterministic; the NCCL result is what
keeps them equal), and a rank whose shard of a tail batch is empty must neither hang nor diverge."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import ROOT  # noqa: F401  (puts the repo root on sys.path, also in spawned workers)
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _build(precision, seed):
    import sunerf_b200 as s
    torch.manual_seed(seed)
    cfg = {'precision': precision} if precision == 'bf16' else {'precision': 'fp32', 'd_filter': 128, 'n_layers': 4}
    return s.EmissionRadiativeTransfer(Rs_per_ds=1, model_config=cfg)


def _batches(n, steps):
    import sunerf_b200 as s
    out = []
    for i in range(steps):
        b = s.rays.synthetic_rays(n, seed=80 + i, H=64, W=64, plate_arcsec=38.0)
        b['t_rand'] = torch.rand(n, 64, generator=torch.Generator().manual_seed(90 + i))
        out.append(b)
    return out


def _worker(rank, world, port, out_dir, precision, n, steps, tail):
    os.environ.update({'RANK': str(rank), 'WORLD_SIZE': str(world), 'LOCAL_RANK': str(rank), 'MASTER_ADDR': '127.0.0.1',
                       'MASTER_PORT': str(port)})
    import sunerf_b200 as s
    from sunerf_b200 import parallel
    torch.cuda.set_device(rank)
    parallel.init_distributed('nccl')
    # rank 1 is seeded differently on purpose: the constructor's broadcast must make the replicas identical
    r = _build(precision, 3 if rank == 0 else 4).cuda()
    tr = s.RayTrainer(r, lr=1e-3)
    losses = []
    for b in _batches(n, steps):
        mine = {k: v.cuda() for k, v in parallel.shard_batch(b, rank, world).items()}
        res = tr.step(mine['rays_o'], mine['rays_d'], mine['times'], mine['target'], t_rand=mine['t_rand'])
        losses.append(res['losses'][0].item())
    if tail:       # a 1-ray global batch: rank 0 gets it, rank 1 gets an empty shard
        b = _batches(1, 1)[0]
        sl = parallel.shard_rows(1, rank, world)
        mine = {k: v[sl].cuda() for k, v in b.items()}
        tr.step(mine['rays_o'], mine['rays_d'], mine['times'], mine['target'], t_rand=mine['t_rand'])
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f'flat_{rank}.npy'), tr.flat.detach().cpu().numpy())
    np.save(os.path.join(out_dir, f'grad_{rank}.npy'), tr.flat_grad.detach().cpu().numpy())
    np.save(os.path.join(out_dir, f'loss_{rank}.npy'), np.array(losses))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def _need2():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')


def test_world2_step_equals_single_process_full_batch(tmp_path):
    _need2()
    import sunerf_b200 as s
    n, steps = 128, 3
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), 'fp32', n, steps, False), nprocs=2, join=True)
    r = _build('fp32', 3).cuda()
    tr = s.RayTrainer(r, lr=1e-3)
    for b in _batches(n, steps):
        tr.step(*(b[k].cuda() for k in ('rays_o', 'rays_d', 'times', 'target')), t_rand=b['t_rand'].cuda())
    flat0, flat1 = np.load(tmp_path / 'flat_0.npy'), np.load(tmp_path / 'flat_1.npy')
    g0 = np.load(tmp_path / 'grad_0.npy')
    assert np.array_equal(flat0, flat1)                               # replicas identical (after a deliberately different seed)
    ref_g, ref_p = tr.flat_grad.cpu().numpy(), tr.flat.cpu().numpy()
    # flat_grad on each rank holds the all-reduced SUM of two half-batch means = 2 x the full-batch mean gradient
    rel = np.linalg.norm(g0 / 2 - ref_g) / np.linalg.norm(ref_g)
    assert rel <= 1e-5, rel
    # 3 Adam steps of lr 1e-3: parameters equal up to the 1/(|g|+eps) amplification of that rounding
    assert np.abs(flat0 - ref_p).max() <= 2e-5, np.abs(flat0 - ref_p).max()


def test_world2_replicas_stay_identical_in_tensor_core_mode_and_survive_an_empty_shard(tmp_path):
    _need2()
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), 'bf16', 256, 4, True), nprocs=2, join=True)
    flat0, flat1 = np.load(tmp_path / 'flat_0.npy'), np.load(tmp_path / 'flat_1.npy')
    assert np.isfinite(flat0).all()
    assert np.array_equal(flat0, flat1)
    assert np.array_equal(np.load(tmp_path / 'grad_0.npy'), np.load(tmp_path / 'grad_1.npy'))


@pytest.mark.timeout(300)
def test_single_process_two_device_render():
    """The reference's render path is ONE process over the visible GPUs: `nn.DataParallel(rendering)` driven from a
    ThreadPoolExecutor (sunerf/evaluation/loader.py:37-39, 143-144, 226-229).  The opt-in shared-memory attributes and the SM
    count are per-device state; the tensor-core modes on device 1 of the same process must give the bits device 0 gives
    (each ray is independent of its neighbours in the batch, so the split does not change a value).
      * `nn.DataParallel(rendering)` itself, one call at a time (from a worker thread, as the loader's pool would run it);
      * `ReplicatedRendering` - resident per-device copies, no per-call broadcast - from a pool of 4 threads at once.
    torch's own `replicate` broadcasts the parameters through NCCL on every forward and deadlocks when two threads enter it
    together (torch 2.11 / NCCL 2.28, stock modules included; tools/dp_diag.py shows the stacks), which is why the
    concurrent half of the reference's pattern is tested on the replacement, not on nn.DataParallel."""
    _need2()
    from concurrent.futures import ThreadPoolExecutor
    import sunerf_b200 as s
    for precision in ('bf16', 'x3'):
        torch.manual_seed(3)
        r = s.EmissionRadiativeTransfer(Rs_per_ds=1, model_config={'precision': precision}).cuda(0)
        r.sampler.perturb = False
        batches = [{k: v.cuda(0) for k, v in s.rays.synthetic_rays(515 + 128 * i, seed=40 + i).items()} for i in range(4)]

        def render(b, module):
            with torch.no_grad():
                return module(b['rays_o'], b['rays_d'], b['times'])['fine_image'].cpu()

        alone = [render(b, r) for b in batches]
        dp = torch.nn.DataParallel(r, device_ids=[0, 1])
        with ThreadPoolExecutor(max_workers=1) as ex:
            one_at_a_time = list(ex.map(lambda b: render(b, dp), batches))
        rep = s.ReplicatedRendering(r, device_ids=[0, 1])
        with ThreadPoolExecutor(max_workers=4) as ex:
            together = list(ex.map(lambda b: render(b, rep), batches * 3))
        for i, img in enumerate(one_at_a_time):
            assert torch.equal(img, alone[i]), (precision, 'DataParallel', i)
        for i, img in enumerate(together):
            assert torch.equal(img, alone[i % len(batches)]), (precision, 'ReplicatedRendering', i)
        # the replicas follow the module after sync_weights()
        with torch.no_grad():
            for p_ in r.parameters():
                p_.mul_(1.01)
        rep.sync_weights()
        assert torch.equal(render(batches[0], rep), render(batches[0], r))
