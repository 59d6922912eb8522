"""Small driver for ncu: a few tensor-core train steps (1024 rays) + forward-only renders (4096 rays).
usage: prof_step.py <steps> <both|train|render> [bf16|x3]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sunerf_b200 as s
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mode = sys.argv[2] if len(sys.argv) > 2 else 'both'
dev = torch.device('cuda', 0)
torch.manual_seed(7)
prec = sys.argv[3] if len(sys.argv) > 3 else 'bf16'
rend = s.EmissionRadiativeTransfer(Rs_per_ds=1, model_config={'precision': prec}).to(dev)
tr = s.RayTrainer(rend)
b = {k: v.to(dev) for k, v in s.rays.synthetic_rays(1024, seed=0).items()}
rb = {k: v.to(dev) for k, v in s.rays.synthetic_rays(4096, seed=1).items()}
for i in range(steps):
    if mode in ('both', 'train'):
        tr.step(b['rays_o'], b['rays_d'], b['times'], b['target'])
    if mode in ('both', 'render'):
        with torch.no_grad():
            rend(rb['rays_o'], rb['rays_d'], rb['times'])
torch.cuda.synchronize()
print('done', s.ops.launch_count())
