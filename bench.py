#!/usr/bin/env python
"""bench.py - SuNeRF ray-render hot path on B200: train rays/s (fwd+bwd+optimiser) and render Msamples/s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU PyTorch path (oracle port)

Workload (BASELINE.json configs[1], `emission_2012_08-193.yaml`): emission SuNeRF, two 8x512 sine MLPs, 64 coarse +
128 hierarchical samples per ray, 1024 synthetic rays per GPU per step, random-init weights.
One "step" = one full training step: sampling, both field networks forward, compositing, asinh-MSE + regulariser
loss, analytic backward into all 3.77 M parameters, (N>1: NCCL all-reduce of the flat gradient), clip + Adam.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAYS_PER_GPU = 1024
S_COARSE, S_FINE = 64, 192
FLOP_FWD_POINT = 3758080          # 2*(84*512 + 7*512^2 + 512*2)          SURVEY.md section 8d
FLOP_BWD_POINT = 7430144          # wgrad 9 layers + dgrad 8 layers
FLOP_TRAIN_RAY = (S_COARSE + S_FINE) * (FLOP_FWD_POINT + FLOP_BWD_POINT)   # 2.8642e9
FLOP_RENDER_RAY = (S_COARSE + S_FINE) * FLOP_FWD_POINT                     # 9.6207e8
RENDER_BATCH = 4096               # render_mhd.yaml batch_size


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'tflops': d.get('bf16_tflops_sustained', 1400.0), 'tflops_burst': d.get('bf16_tflops', 1590.0),
                'hbm': d.get('hbm_gbs', 6650.0), 'src': 'measured'}
    return {'tflops': 1400.0, 'tflops_burst': 1590.0, 'hbm': 6650.0, 'src': 'fallback'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms; started before the warm-up (nvidia-smi takes a moment
    to come up), only the samples that fall inside the timed window [mark_start, mark_end] are reported."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(',')]))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        t0, t1 = self.t0 or 0.0, (self.t1 or time.time()) + 0.15
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 and len(r) >= 6]
        if not rows:                       # window shorter than the sampling period: take the nearest samples
            rows = [r for ts, r in self.rows if len(r) >= 6][-3:]
        sm = [float(r[0]) for r in rows if r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in rows for n, v in zip(names, r[2:6]) if v.lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


def synthetic_batch(n, seed):
    """SURVEY.md section 8d synthetic rays: 1 AU observers, helioprojective pixel directions, t~U(0,30 d), target~U(0,1)."""
    from sunerf_b200.rays import synthetic_rays
    return synthetic_rays(n, seed=seed, H=256, W=256, plate_arcsec=9.4, t_days=30.0)


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's own CPU PyTorch path for the same step (oracle port: same ATen ops as the reference, which
    cannot travel to the GPU box), all host threads.  Each step is the FULL 1024-ray step of the config (BASELINE.md
    section 3) unless --ref-rays says otherwise; the line states the ray count it ran."""
    if rank != 0:
        return
    from oracle import sunerf_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = args.ref_rays
    b = synthetic_batch(n, seed=0)
    torch.manual_seed(7)
    pc, pf = orc.FieldParams.init(1).requires_grad_(), orc.FieldParams.init(2).requires_grad_()
    opt = orc.AdamState(pc.tensors() + pf.tensors())
    cfg = orc.RenderConfig(kind='emission')
    gen = torch.Generator().manual_seed(3)

    def step():
        t_rand = torch.rand(n, 64, generator=gen)
        return orc.train_step(cfg, pc, pf, opt, b['rays_o'], b['rays_d'], b['times'], b['target'], None, t_rand)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    line = {'metric': 'train_rays_per_s', 'value': v, 'unit': 'rays/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
            'config': workload_config(args.gpus),
            'cpu_baseline': {'value': v, 'unit': 'rays/s', 'cores': cores, 'kind': 'port', 'rays_per_step': n,
                             'sample': (f'the full {n}-ray train step' if n == RAYS_PER_GPU else f'{n} of the {RAYS_PER_GPU} rays per step')
                                       + f' x {args.steps} steps after {args.warmup} warm-up, torch CPU fp32 oracle, {cores} threads'},
            'e2e': {'value': v, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def workload_config(gpus):
    return {'workload': 'emission_2012_08-193.yaml: emission SuNeRF train step, 2 x (84->512x8->2) sine MLP, 64+128 samples/ray',
            'rays_per_gpu': RAYS_PER_GPU, 'global_rays': RAYS_PER_GPU * gpus, 'samples_per_ray': S_COARSE + S_FINE,
            'parallelism': f'ray-shard dp{gpus}, one NCCL all-reduce of the flat fp32 gradient per step',
            'launch': 'whole step captured once and replayed as a CUDA graph (RayTrainer(use_cuda_graph=True)); the coarse '
                      'network\'s backward runs on a side stream beside the fine one (roofline pass: serial)',
            'l2': 'per-step working set (saved layer activations, >1 GB) exceeds the 126 MB L2; no explicit flush'}


def cpu_baseline(sample_rays=RAYS_PER_GPU, render_rays=RENDER_BATCH):
    """1 warm-up + best of 3 of the full train step at the config's ray count, and of a forward-only
    render batch of 4096 rays, on all host cores (about 25 s of CPU work on the 16-core box)."""
    from oracle import sunerf_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b = synthetic_batch(sample_rays, seed=0)
    pc, pf = orc.FieldParams.init(1).requires_grad_(), orc.FieldParams.init(2).requires_grad_()
    opt = orc.AdamState(pc.tensors() + pf.tensors())
    cfg = orc.RenderConfig(kind='emission')
    gen = torch.Generator().manual_seed(3)
    best = None
    for i in range(4):   # 1 warm-up + best of 3
        t_rand = torch.rand(sample_rays, 64, generator=gen)
        t0 = time.perf_counter()
        orc.train_step(cfg, pc, pf, opt, b['rays_o'], b['rays_d'], b['times'], b['target'], None, t_rand)
        dt = time.perf_counter() - t0
        if i > 0:
            best = dt if best is None else min(best, dt)
    # forward-only rendering of the same rays (the `render` metric's CPU counterpart)
    rbest = None
    rb = synthetic_batch(render_rays, seed=1)
    with torch.no_grad():
        for i in range(4):
            t0 = time.perf_counter()
            orc.render(cfg, pc, pf, rb['rays_o'], rb['rays_d'], rb['times'], None, None)
            dt = time.perf_counter() - t0
            if i > 0:
                rbest = dt if rbest is None else min(rbest, dt)
    return {'value': sample_rays / best, 'unit': 'rays/s', 'cores': cores, 'kind': 'port', 'rays_per_step': sample_rays,
            'sample': f'1 warm-up + best of 3 full train steps of {sample_rays} rays (the config size), torch CPU fp32 oracle, {cores} threads',
            'render_Msamples_per_s': render_rays * (S_COARSE + S_FINE) / rbest / 1e6,
            'render_sample': f'1 warm-up + best of 3 forward-only renders of one {render_rays}-ray batch'}


# ------------------------------------------------------------------------------------------ this repo's arm
def teardown(world):
    """Destroy the NCCL communicator AFTER the line is printed; the measurements are complete by then, so a teardown that
    raises or wedges (NCCL waits for every CUDA graph that captured the communicator) must not fail the run."""
    if world <= 1:
        return
    t = threading.Timer(30.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    try:
        torch.distributed.destroy_process_group()
    except Exception as e:
        print(f'[bench] process-group teardown: {e!r}', file=sys.stderr, flush=True)
    t.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('SUNERF_B200_PRECISION', 'bf16'), choices=['fp32', 'x3', 'bf16'])
    ap.add_argument('--ref-rays', type=int, default=RAYS_PER_GPU, help='rays per step of the reference arm (default: the full config batch)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--serial-backward', action='store_true', help='coarse backward after the fine one instead of beside it (A/B aid)')
    ap.add_argument('--quick', action='store_true', help='headline step only (A/B aid): prints {"ms_per_step", "rays_per_s"} and exits')
    ap.add_argument('--no-cuda-graph', action='store_true', help='launch the ~30 kernels of a step one by one instead of replaying a graph')
    args = ap.parse_args()
    if os.environ.get('SNF_BENCH_WATCHDOG'):        # developer aid: dump every thread's stack and exit if the run wedges
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ['SNF_BENCH_WATCHDOG']), exit=True)
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    import sunerf_b200 as s
    from sunerf_b200 import ops
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.distributed.init_process_group('nccl', device_id=dev)
    pk = peaks()

    torch.manual_seed(7)                      # run_density_temperature.py:17; same init on every rank (no broadcast)
    rend = s.EmissionRadiativeTransfer(Rs_per_ds=1, model_config={'precision': args.precision}).to(dev)
    trainer = s.RayTrainer(rend, use_cuda_graph=not args.no_cuda_graph)
    trainer.overlap_backward = not args.serial_backward
    N = RAYS_PER_GPU
    b = synthetic_batch(N * world, seed=0)    # global batch, sharded by rank: rank r owns rays [r*N, (r+1)*N)
    host = {k: v[rank * N:(rank + 1) * N].contiguous().pin_memory() for k, v in b.items()}
    devb = {k: v.to(dev) for k, v in host.items()}
    gen = torch.Generator(device=dev).manual_seed(100 + rank)

    # ---- CUDA-event instrumentation of the dominant kernel group (field-network forward + backward)
    mlp_events = []            # (start, stop, 'fwd' | 'bwd') on torch's current stream = the stream the kernels launch on
    _fwd, _bwd = ops.mlp_forward, ops.mlp_backward

    def timed(fn, tag):
        def w(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = fn(*a, **k); e1.record()
            mlp_events.append((e0, e1, tag))
            return r
        return w

    def step_resident():
        t_rand = torch.rand((N, S_COARSE), device=dev, generator=gen)
        return trainer.step(devb['rays_o'], devb['rays_d'], devb['times'], devb['target'], t_rand=t_rand)

    def step_e2e():                         # the public call with HOST (pinned) buffers: the H2D copies are part of the step
        t_rand = torch.rand((N, S_COARSE), device=dev, generator=gen)
        res = trainer.step(host['rays_o'], host['rays_d'], host['times'], host['target'], t_rand=t_rand)
        return res['losses'].cpu()          # device->host read of the step's result (synchronises)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed_loop(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return ms.item()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step_resident()
    clocks.mark_start()
    l0 = ops.launch_count()
    ms = timed_loop(step_resident, args.steps)          # the headline: the step replayed as one CUDA graph (unless --no-cuda-graph)
    launches = ops.launch_count() - l0
    if args.quick:
        ms2 = timed_loop(step_resident, args.steps)
        if rank == 0:
            print(json.dumps({'ms_per_step': min(ms, ms2) / args.steps, 'rays_per_s': N * world * args.steps / (min(ms, ms2) * 1e-3),
                              'n_gpus': world, 'early_reduce': trainer.early_coarse_reduce, 'reserve_sms': trainer.reserve_sms}), flush=True)
        trainer._graph = None                  # a captured step holds NCCL work: release it before the communicator goes
        del trainer
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        teardown(world)
        return
    # ---- roofline pass: the same K steps launched kernel by kernel, so that CUDA events can bracket the field-network
    #      launches on their stream (events cannot be read back from inside a replayed graph); same kernels, same data
    graphed = trainer.use_cuda_graph
    trainer.use_cuda_graph = False
    trainer.overlap_backward = False                   # one kernel at a time, so that each is timed alone
    ops.mlp_forward, ops.mlp_backward = timed(_fwd, 'fwd'), timed(_bwd, 'bwd')
    for _ in range(2):
        step_resident()
    mlp_events.clear()
    from sunerf_b200 import _lib as _snf_lib
    import ctypes as _ct
    if args.precision == 'bf16':
        _snf_lib.lib().snf_debug_time_backward(1)     # per-kernel events inside snf_mlp_bwd_bf16 (no host sync)
    ms_eager = timed_loop(step_resident, args.steps)
    trainer.use_cuda_graph = graphed
    trainer.overlap_backward = not args.serial_backward
    ops.mlp_forward, ops.mlp_backward = _fwd, _bwd
    mlp_ms = sum(a.elapsed_time(bb) for a, bb, _ in mlp_events) / max(1, args.steps)   # per step: 2 fwd + 2 bwd groups
    fwd_ms = sum(a.elapsed_time(bb) for a, bb, t in mlp_events if t == 'fwd') / max(1, args.steps)
    bwd_ms = mlp_ms - fwd_ms
    kernels = None
    if args.precision == 'bf16':
        buf = (_ct.c_double * 3)()
        ncalls = _snf_lib.lib().snf_debug_backward_ms(buf)
        _snf_lib.lib().snf_debug_time_backward(0)
        if ncalls > 0:
            per_step = [buf[i] / ncalls * 2 for i in range(3)]          # two backward calls (fine + coarse network) per step
            pts = N * (S_COARSE + S_FINE)
            flop = {'mlp_fwd_bf16_kernel<train>': pts * FLOP_FWD_POINT, 'mlp_dgrad_bf16_kernel': pts * 7 * 2 * 512 * 512,
                    'mlp_wgrad_bf16_kernel': pts * (84 + 7 * 512) * 512 * 2}
            kms = {'mlp_fwd_bf16_kernel<train>': fwd_ms, 'mlp_dgrad_bf16_kernel': per_step[0], 'mlp_wgrad_bf16_kernel': per_step[1]}
            kernels = [{'kernel': k, 'launches_per_step': 2, 'ms_per_step': kms[k], 'algorithmic_flop_per_step': flop[k],
                        'achieved_tflops': flop[k] / (kms[k] * 1e-3) / 1e12, 'frac': flop[k] / (kms[k] * 1e-3) / 1e12 / pk['tflops']}
                       for k in kms]
            kernels.append({'kernel': 'out_wgrad_bf16_kernel', 'launches_per_step': 2, 'ms_per_step': per_step[2], 'bound': 'hbm',
                            'algorithmic_bytes_per_step': pts * 1024, 'achieved_GBs': pts * 1024 / (per_step[2] * 1e-3) / 1e9})
    for _ in range(2):
        step_e2e()
    ms_e2e = timed_loop(step_e2e, args.steps)
    clocks.mark_end()
    clk = clocks.stop() if rank == 0 else None
    trainer.check_finite()

    # ---- forward-only render throughput (render_mhd.yaml batch of 4096 rays, hierarchical sampling on)
    rb = synthetic_batch(RENDER_BATCH, seed=1)
    rdev = {k: v.to(dev) for k, v in rb.items()}

    def render_once():
        with torch.no_grad():
            return rend(rdev['rays_o'], rdev['rays_d'], rdev['times'])
    for _ in range(3):
        render_once()
    ms_render = timed_loop(render_once, max(3, args.steps // 2)) / max(3, args.steps // 2)

    # ---- full-image novel view (render_mhd.yaml at 1024^2): pose -> rays generated on the device ->
    #      batched render -> image assembled on the device; this rank's share of the rows, no collective
    renderer = s.ObserverRenderer(rend, (1024, 1024), plate_arcsec=2.4)
    rows = s.parallel.shard_rows(1024, rank, world)

    def image_once():
        return renderer.render_observer_image(0.05, 1.0, 3.0, batch_size=RENDER_BATCH, rows=rows, as_numpy=False)
    image_once()                 # warm-up at full size: the first call pays cudaMalloc for ~1.5 GB of output maps
    ms_image = timed_loop(image_once, 1)

    # ---- the exact mode on tensor cores (precision 'x3': fp32-mode gates, three fp16 MMAs per product) on the same workload
    exact = None
    if args.precision == 'bf16':
        torch.manual_seed(7)
        rx = s.EmissionRadiativeTransfer(Rs_per_ds=1, model_config={'precision': 'x3'}).to(dev)
        tx = s.RayTrainer(rx, use_cuda_graph=not args.no_cuda_graph)

        def step_x3():
            return tx.step(devb['rays_o'], devb['rays_d'], devb['times'], devb['target'],
                           t_rand=torch.rand((N, S_COARSE), device=dev, generator=gen))
        for _ in range(4):
            step_x3()
        nx = max(5, args.steps // 2)
        ms_x = timed_loop(step_x3, nx) / nx
        tx.check_finite()
        tx._graph = None
        del tx, step_x3

        def render_x3():
            with torch.no_grad():
                return rx(rdev['rays_o'], rdev['rays_d'], rdev['times'])
        for _ in range(2):
            render_x3()
        ms_rx = timed_loop(render_x3, 5) / 5
        exact = {'precision': 'x3 (split-precision fp16 pairs on tcgen05; gates 1e-5 intensities / 1e-3 gradients)',
                 'train_rays_per_s': N * world / (ms_x * 1e-3), 'ms_per_step': ms_x,
                 'tensor_flop_frac': N * (S_COARSE + S_FINE) * (3 * FLOP_FWD_POINT + 7 * 2 * 512 * 512 * 2 + (84 + 7 * 512) * 512 * 2)
                 / (ms_x * 1e-3) / 1e12 / pk['tflops'],
                 'render_Msamples_per_s': RENDER_BATCH * (S_COARSE + S_FINE) * world / (ms_rx * 1e-3) / 1e6, 'ms_per_render_batch': ms_rx}
        del rx
        torch.cuda.empty_cache()

    # ---- render_mhd.yaml as shipped (BASELINE.json configs[4]): density-temperature head, C = 6, pixel_intensity_factor
    #      1e10 (image_render.py:267), field = SimpleStar (render_mhd.yaml:1) and a trained-size NeRF_DT: a 4096-ray batch
    #      and the 1024^2 image each
    wl6 = torch.tensor([94., 171., 193., 211., 304., 335.], device=dev)
    mhd = {}
    for field in ('SimpleStar', 'NeRF_DT'):
        torch.manual_seed(7)
        kw = {'model': s.SimpleStar} if field == 'SimpleStar' else {'model': s.NeRF_DT, 'model_config': {'precision': args.precision}}
        rmhd = s.DensityTemperatureRadiativeTransfer(Rs_per_ds=1, pixel_intensity_factor=1e10, **kw).to(dev)
        wlb = wl6[None].expand(RENDER_BATCH, 6).contiguous()

        def render_mhd_once():
            with torch.no_grad():
                return rmhd(rdev['rays_o'], rdev['rays_d'], rdev['times'], wlb)
        for _ in range(3):
            render_mhd_once()
        nrep = max(3, args.steps // 2)
        ms_b = timed_loop(render_mhd_once, nrep) / nrep
        ren = s.ObserverRenderer(rmhd, (1024, 1024), plate_arcsec=2.4)
        image_mhd = lambda: ren.render_observer_image(0.05, 1.0, 3.0, wl=wl6.cpu().numpy(), batch_size=RENDER_BATCH, rows=rows,
                                                      as_numpy=False)
        image_mhd()              # warm-up at full size (allocations)
        ms_i = timed_loop(image_mhd, 1)
        mhd[field] = {'batch_4096': {'ms': ms_b, 'Msamples_per_s': RENDER_BATCH * (S_COARSE + S_FINE) * world / (ms_b * 1e-3) / 1e6},
                      'image_1024': {'ms': ms_i, 'Msamples_per_s': 1024 * 1024 * (S_COARSE + S_FINE) / (ms_i * 1e-3) / 1e6}}
        del rmhd, ren

    # ---- density-temperature config (DT_2012_11.yaml: NeRF_DT + AIA response head, 3072 rays, 7 channels, half of the
    #      rays with the STEREO channel mask) - reported next to the headline, same fast path
    trainer._graph = None                  # closures above still hold the trainer: drop its captured step explicitly
    del trainer, rend, renderer
    torch.cuda.empty_cache()
    torch.manual_seed(7)
    rend_dt = s.DensityTemperatureRadiativeTransfer(Rs_per_ds=1, model=s.NeRF_DT, pixel_intensity_factor=1e17,
                                                    model_config={'precision': args.precision}).to(dev)
    tr_dt = s.RayTrainer(rend_dt)
    Nd = 3072
    bd = synthetic_batch(Nd, seed=2)
    wl = torch.tensor([94., 131., 171., 193., 211., 304., 335.]).repeat(Nd, 1)
    wl[Nd // 2:] = torch.tensor([0., 0., 171., 193., 211., 304., 0.])            # multi_thermal_loader.py:162-168
    dd = {k: v.to(dev) for k, v in bd.items()}
    dd['wl'] = wl.to(dev)
    dd['target'] = torch.rand(Nd, 7, device=dev)

    def step_dt():
        return tr_dt.step(dd['rays_o'], dd['rays_d'], dd['times'], dd['target'], dd['wl'])
    for _ in range(3):
        step_dt()
    ms_dt = timed_loop(step_dt, 5) / 5
    tr_dt.check_finite()

    # a captured step holds NCCL work: release every graph before the communicator goes
    tr_dt._graph = None
    del tr_dt, step_dt, step_resident, step_e2e
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    if rank != 0:
        teardown(world)
        return
    value = N * world * args.steps / (ms * 1e-3)
    e2e = N * world * args.steps / (ms_e2e * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    achieved = N * FLOP_TRAIN_RAY / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
    # DRAM bytes of the field-network kernels of one step: ncu cannot run inside a timed bench, so this is the committed
    # `ncu --set full` capture of this same command (profiles/r02_ncu_traffic.json, falling back to round 1's)
    traffic, traffic_src = None, None
    for name in ('r02_ncu_traffic.json', 'r01_ncu_traffic.json'):
        tpath = os.path.join(ROOT, 'profiles', name)
        if args.precision == 'bf16' and os.path.exists(tpath):
            traffic, traffic_src = json.load(open(tpath)).get('train_step_mlp_traffic_bytes'), 'profiles/' + name
            break
    step_tflops = N * FLOP_TRAIN_RAY / (ms / args.steps * 1e-3) / 1e12
    line = {'metric': 'train_rays_per_s', 'value': value, 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': {'bf16': 'f16 operands (weights, activations, scaled gradients), f32 accumulate',
                                           'x3': 'f16 (hi, lo) operand pairs, 3 MMAs per product, f32 accumulate', 'fp32': 'f32'}[args.precision],
            'data': 'synthetic',
            'config': workload_config(world),
            'e2e': {'value': e2e, 'unit': 'rays/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 16,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(launches) * world,
            'roofline': {'bound': 'tensor', 'kernel': f'field-network MLP forward+backward ({args.precision})',
                         'achieved': achieved, 'peak': pk['tflops'], 'unit': 'TFLOP/s', 'frac': achieved / pk['tflops'],
                         'traffic': traffic, 'traffic_source': traffic_src,
                         # the same FLOPs over the HEADLINE step time (graph replay, coarse backward overlapped, every other
                         # kernel of the step included): the self-consistent whole-step figure
                         'frac_from_step': step_tflops / pk['tflops'], 'achieved_from_step': step_tflops,
                         'peak_source': pk['src'] + ' (sustained cuBLAS bf16: the kernels run inside a long step)',
                         'launches_per_step': 6 if args.precision == 'bf16' else None, 'kernels': kernels,
                         'forward': {'ms_per_step': fwd_ms, 'tflops': N * (S_COARSE + S_FINE) * FLOP_FWD_POINT / (fwd_ms * 1e-3) / 1e12},
                         'backward': {'ms_per_step': bwd_ms, 'tflops': N * (S_COARSE + S_FINE) * FLOP_BWD_POINT / (bwd_ms * 1e-3) / 1e12},
                         'mlp_ms_per_step': mlp_ms, 'mlp_share_of_step': mlp_ms / (ms_eager / args.steps),
                         'timing': 'CUDA events around the field-network launches (torch current stream = launch stream) in an '
                                   'eager pass of the same K steps run right after the timed region; the timed region itself '
                                   + ('replays the step as one CUDA graph' if graphed else 'is launched eagerly too'),
                         'eager_ms_per_step': ms_eager / args.steps,
                         'algorithmic_flop_per_step': N * FLOP_TRAIN_RAY},
            'render': {'metric': 'render_Msamples_per_s', 'value': RENDER_BATCH * (S_COARSE + S_FINE) * world / (ms_render * 1e-3) / 1e6,
                       'unit': 'Msamples/s', 'rays_per_batch': RENDER_BATCH, 'ms_per_batch': ms_render,
                       'tensor_frac': RENDER_BATCH * FLOP_RENDER_RAY / (ms_render * 1e-3) / 1e12 / pk['tflops_burst'],
                       'tensor_peak': pk['tflops_burst'], 'tensor_peak_source': pk['src'] + ' (burst cuBLAS bf16: a 3 ms forward timed alone)',
                       'image_1024': {'ms': ms_image, 'Msamples_per_s': 1024 * 1024 * (S_COARSE + S_FINE) / (ms_image * 1e-3) / 1e6,
                                      'what': 'ObserverRenderer.render_observer_image, 1024x1024 pixels, rays generated on the '
                                              'device, rows sharded over ranks, no collective'}},
            'exact_mode': exact,
            'render_mhd': {'workload': 'render_mhd.yaml: density-temperature head, C=6, F=1e10, hierarchical sampling on; per field a '
                                       '4096-ray batch (ms per batch) and ObserverRenderer.render_observer_image at 1024x1024',
                           **mhd},
            'dt_train': {'workload': 'DT_2012_11.yaml: density-temperature SuNeRF train step, 3072 rays/GPU, C=7 (half STEREO-masked)',
                         'rays_per_s': Nd * world / (ms_dt * 1e-3), 'ms_per_step': ms_dt},
            'clocks': clk}
    if not args.no_cpu_baseline and world == 1:
        line['cpu_baseline'] = cpu_baseline()
    print(json.dumps(line), flush=True)
    teardown(world)


if __name__ == '__main__':
    main()
