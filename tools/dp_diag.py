"""Step-by-step diagnosis of single-process two-device use (run with 2 GPUs): prints progress, dumps every thread's stack
and exits if a step wedges."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(int(os.environ.get('DP_DIAG_LIMIT', '50')), exit=True)
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sunerf_b200 as s
from concurrent.futures import ThreadPoolExecutor

def log(*a):
    print(f'[{time.time() - T0:6.2f}s]', *a, flush=True)

T0 = time.time()
prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
torch.manual_seed(3)
r = s.EmissionRadiativeTransfer(Rs_per_ds=1, model_config={'precision': prec}).cuda(0)
r.sampler.perturb = False
b = {k: v.cuda(0) for k, v in s.rays.synthetic_rays(515, seed=40).items()}
with torch.no_grad():
    ref = r(b['rays_o'], b['rays_d'], b['times'])['fine_image'].cpu()
log('device 0 render ok')
import copy
r1 = copy.deepcopy(r).cuda(1)
b1 = {k: v.cuda(1) for k, v in b.items()}
with torch.cuda.device(1), torch.no_grad():
    o1 = r1(b1['rays_o'], b1['rays_d'], b1['times'])['fine_image']
    torch.cuda.synchronize(1)
log('device 1 render ok, equal:', torch.equal(o1.cpu(), ref))
dp = torch.nn.DataParallel(r, device_ids=[0, 1])
with torch.no_grad():
    o = dp(b['rays_o'], b['rays_d'], b['times'])['fine_image']
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
log('DataParallel forward (main thread) ok, equal:', torch.equal(o.cpu(), ref))

def render(_):
    with torch.no_grad():
        return dp(b['rays_o'], b['rays_d'], b['times'])['fine_image'].cpu()
with ThreadPoolExecutor(max_workers=1) as ex:
    outs = list(ex.map(render, range(2)))
log('DataParallel from 1 worker thread ok, equal:', all(torch.equal(o, ref) for o in outs))
with ThreadPoolExecutor(max_workers=3) as ex:
    outs = list(ex.map(render, range(6)))
log('DataParallel from 3 worker threads ok, equal:', all(torch.equal(o, ref) for o in outs))
