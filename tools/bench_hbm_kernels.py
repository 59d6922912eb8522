"""Stand-alone HBM roofline of the sampling / compositing kernels (N >= 2^20 rays so the working
set exceeds the 126 MB L2).  achieved = ALGORITHMIC bytes per launch / CUDA-event time; peak = MEASURED_PEAKS.json.

    python tools/bench_hbm_kernels.py [--rays 1048576] [--json profiles/xxx.json]
"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sunerf_b200 as s
from sunerf_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument('--rays', type=int, default=1 << 20)
ap.add_argument('--iters', type=int, default=10)
ap.add_argument('--json', default=None)
ap.add_argument('--only', default=None)
ap.add_argument('--dt-channels', default='7', help='comma-separated channel counts for the K6 rows (default: 7, DT_2012_11.yaml)')
args = ap.parse_args()
dev = torch.device('cuda', 0)
N = args.rays
pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json'))) \
    if os.path.exists('MEASURED_PEAKS.json') else {'hbm_gbs': 6650.0}
peak = pk['hbm_gbs']
g = torch.Generator(device=dev).manual_seed(0)
base = s.rays.synthetic_rays(4096, seed=0)
rep = N // 4096
rays_o = base['rays_o'].to(dev).repeat(rep, 1).contiguous()
rays_d = base['rays_d'].to(dev).repeat(rep, 1).contiguous()
t_vals = torch.linspace(0, 1, 64, device=dev)
u = torch.linspace(0, 1, 128, device=dev)
t_rand = torch.rand(N, 64, device=dev, generator=g)


def timeit(fn, nbytes, name):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    gbs = nbytes / (ms * 1e-3) / 1e9
    row = {'kernel': name, 'rays': N, 'ms': ms, 'algorithmic_bytes': nbytes, 'achieved_GBs': gbs, 'peak_GBs': peak, 'frac': gbs / peak}
    print(json.dumps(row), flush=True)
    return row


rows = []
want = lambda k: args.only is None or args.only in k
# a1 stratified: in o,d (24 B) + t_rand (256 B) ; out z (256 B)  -> 536 B/ray
if want('stratified'):
    rows.append(timeit(lambda: ops.stratified_sample(rays_o, rays_d, t_vals, t_rand, 1.3, 1.0), N * 536, 'K1 stratified_sample (S=64)'))
z64, _ = ops.stratified_sample(rays_o, rays_d, t_vals, t_rand, 1.3, 1.0)
w64 = torch.rand(N, 64, device=dev, generator=g)
# a2 resampler: in z (256) + w (256) ; out new_z (512) + z_comb (768) -> 1792 B/ray
if want('hier'):
    rows.append(timeit(lambda: ops.hier_resample(z64, w64, u), N * 1792, 'K2 hier_resample (64 -> +128)'))
_, z192, _, _ = ops.hier_resample(z64, w64, u)
del w64, t_rand
for S, z in ((64, z64), (192, z192)):
    raw = torch.randn(N, S, 2, device=dev, generator=g) * 0.5
    # a8 forward: raw 8S + z 4S in, weights 4S + absorption 4S + image 4 out, d 12 -> 20 S + 16
    if want('emission_fwd'):
        rows.append(timeit(lambda: ops.composite_emission_fwd(raw, z, rays_d), N * (20 * S + 16), f'K5 composite_emission_fwd (S={S})'))
    g_img = torch.randn(N, 1, device=dev, generator=g)
    g_abs = torch.randn(N, S, device=dev, generator=g) if S == 192 else None
    # backward: raw 8S + z 4S (+ g_abs 4S) in, g_raw 8S out, d 12 + g_image 4 -> 24 S + 16 (20 S + 16 without g_abs)
    if want('emission_bwd'):
        nb = N * ((24 if g_abs is not None else 20) * S + 16)
        rows.append(timeit(lambda: ops.composite_emission_bwd(raw, z, rays_d, g_img, g_abs), nb, f'K5 composite_emission_bwd (S={S})'))
    del raw, g_img, g_abs
# a9 DT head, C = 7 (DT_2012_11.yaml), S = 192, N/4 rays (the [N,S,2] tensors are the same size as above)
if want('dt'):
    from sunerf_b200 import rendering as R
    Nd = N // 2
    rend = s.DensityTemperatureRadiativeTransfer(Rs_per_ds=1).to(dev)
    wl = torch.tensor([94., 131., 171., 193., 211., 304., 335.], device=dev).repeat(Nd, 1).contiguous()
    inf = torch.randn(Nd, 192, 2, device=dev, generator=g) * 0.3 + torch.tensor([10.0, 5.5], device=dev)
    zz = z192[:Nd].contiguous()
    la = torch.full((7,), 1e-6, device=dev)
    vc = torch.ones(1, device=dev)
    tx, ty = rend._table_x, rend._table_y
    for C in [int(c) for c in args.dt_channels.split(',')]:
        wlc = wl[:, :C].contiguous()
        rows.append(timeit(lambda: ops.composite_dt_fwd(inf, zz, wlc, la, vc, tx, ty, 1e17), Nd * (20 * 192 + 8 * C), f'K6 composite_dt_fwd (S=192, C={C})'))
        gi = torch.randn(Nd, C, device=dev, generator=g)
        rows.append(timeit(lambda: ops.composite_dt_bwd(inf, zz, wlc, la, vc, tx, ty, 1e17, gi), Nd * (28 * 192 + 8 * C), f'K6 composite_dt_bwd (S=192, C={C})'))
if args.json:
    json.dump(rows, open(args.json, 'w'), indent=1)
