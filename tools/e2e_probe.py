"""Where the end-to-end step (pinned host rays in, loss read back) loses time against the device-resident graph replay.
usage: python tools/e2e_probe.py [steps]"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sunerf_b200 as s

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device('cuda', 0)
torch.manual_seed(7)
rend = s.EmissionRadiativeTransfer(Rs_per_ds=1, model_config={'precision': 'bf16'}).to(dev)
tr = s.RayTrainer(rend, use_cuda_graph=True)
N = 1024
b = s.rays.synthetic_rays(N, seed=0, H=256, W=256, plate_arcsec=9.4, t_days=30.0)
host = {k: v.contiguous().pin_memory() for k, v in b.items()}
devb = {k: v.to(dev) for k, v in host.items()}
gen = torch.Generator(device=dev).manual_seed(100)
t_fixed = torch.rand((N, 64), device=dev, generator=gen)


def loop(fn, n=steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) / n * 1e3


def resident():
    return tr.step(devb['rays_o'], devb['rays_d'], devb['times'], devb['target'], t_rand=torch.rand((N, 64), device=dev, generator=gen))


def resident_fixed_rand():
    return tr.step(devb['rays_o'], devb['rays_d'], devb['times'], devb['target'], t_rand=t_fixed)


def host_in_no_readback():
    return tr.step(host['rays_o'], host['rays_d'], host['times'], host['target'], t_rand=torch.rand((N, 64), device=dev, generator=gen))


def resident_readback():
    return resident()['losses'].cpu()


def e2e():
    return host_in_no_readback()['losses'].cpu()


pin = torch.empty(4, dtype=torch.float32).pin_memory()


def e2e_pinned_readback():
    r = host_in_no_readback()
    pin.copy_(r['losses'], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return pin


for _ in range(6):
    resident()
for name, fn in (('resident', resident), ('resident_fixed_rand', resident_fixed_rand), ('host_in_no_readback', host_in_no_readback),
                 ('resident_readback', resident_readback), ('e2e', e2e), ('e2e_pinned_readback', e2e_pinned_readback),
                 ('resident', resident), ('e2e', e2e)):
    g, w = loop(fn)
    print(f'{name:24s} gpu {g:.3f} ms/step   wall {w:.3f} ms/step', flush=True)

# host cost of one step() call with nothing to wait for (CPU time until the launch returns)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    host_in_no_readback()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f'host time per step() call (async): {(t1 - t0) / steps * 1e3:.3f} ms')
