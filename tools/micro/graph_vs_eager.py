import sys, time, torch
sys.path.insert(0,'/root/repo')
import sunerf_b200 as s
dev=torch.device('cuda',0)
for graph in (False, True):
    torch.manual_seed(7)
    rend=s.EmissionRadiativeTransfer(Rs_per_ds=1, model_config={'precision':'bf16'}).to(dev)
    tr=s.RayTrainer(rend, use_cuda_graph=graph)
    b={k:v.to(dev) for k,v in s.rays.synthetic_rays(1024, seed=0).items()}
    host={k:v.pin_memory() for k,v in s.rays.synthetic_rays(1024, seed=0).items()}
    gen=torch.Generator(device=dev).manual_seed(1)
    def step():
        return tr.step(b['rays_o'],b['rays_d'],b['times'],b['target'],t_rand=torch.rand((1024,64),device=dev,generator=gen))
    def step_e2e():
        d={k:v.to(dev,non_blocking=True) for k,v in host.items()}
        return tr.step(d['rays_o'],d['rays_d'],d['times'],d['target'],t_rand=torch.rand((1024,64),device=dev,generator=gen))['losses'].cpu()
    for _ in range(6): step()
    for fn,name in ((step,'resident'),(step_e2e,'e2e')):
        torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40): fn()
        e1.record(); torch.cuda.synchronize()
        print('graph' if graph else 'eager', name, round(e0.elapsed_time(e1)/40,4),'ms/step')
