"""Quick timing of the x3 (exact, tensor-core) mode: train step at 1024 rays, render of a 4096-ray batch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sunerf_b200 as s
prec = sys.argv[1] if len(sys.argv) > 1 else 'x3'
dev = torch.device('cuda', 0)
torch.manual_seed(7)
rend = s.EmissionRadiativeTransfer(Rs_per_ds=1, model_config={'precision': prec}).to(dev)
tr = s.RayTrainer(rend, use_cuda_graph=True)
b = {k: v.to(dev) for k, v in s.rays.synthetic_rays(1024, seed=0).items()}
rb = {k: v.to(dev) for k, v in s.rays.synthetic_rays(4096, seed=1).items()}
def timeit(fn, n):
    for _ in range(4): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t = timeit(lambda: tr.step(b['rays_o'], b['rays_d'], b['times'], b['target']), 20)
def render():
    with torch.no_grad(): rend(rb['rays_o'], rb['rays_d'], rb['times'])
r = timeit(render, 10)
print(f'{prec}: train {t:.3f} ms/step ({1024 / t:.1f} k rays/s)   render {r:.3f} ms/batch ({4096 * 256 / r / 1e3:.1f} Msamples/s)')
