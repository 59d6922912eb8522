import torch, sys
sys.path.insert(0,'/root/repo')
import sunerf_b200 as s
from sunerf_b200 import ops
dev='cuda'
N=1<<18
base=s.rays.synthetic_rays(4096, seed=0)
ro=base['rays_o'].to(dev).repeat(N//4096,1).contiguous(); rd=base['rays_d'].to(dev).repeat(N//4096,1).contiguous()
g=torch.Generator(device=dev).manual_seed(0)
tv=torch.linspace(0,1,64,device=dev); u=torch.linspace(0,1,128,device=dev)
tr=torch.rand(N,64,device=dev,generator=g)
z,_=ops.stratified_sample(ro,rd,tv,tr,1.3,1.0)
for name,w in (('uniform random weights', torch.rand(N,64,device=dev,generator=g)), ('peaked', torch.rand(N,64,device=dev,generator=g)**8)):
    nz,zc,_,_=ops.hier_resample(z,w,u)
    bad=(nz[:,1:]<nz[:,:-1]).any(1).float().mean().item()
    badz=(z[:,1:]<z[:,:-1]).any(1).float().mean().item()
    print(name,'rays with unsorted new_z:',bad,' unsorted z:',badz, ' z_comb sorted:', bool((zc[:,1:]>=zc[:,:-1]).all()))
