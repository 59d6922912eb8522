"""Developer aid: read the in-kernel cycle counters of the bf16 forward (library built with SNF_NVCC_EXTRA=-DSNF_PROF)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sunerf_b200 as s
from sunerf_b200 import _lib
dev = torch.device('cuda', 0)
torch.manual_seed(7)
rend = s.EmissionRadiativeTransfer(Rs_per_ds=1, model_config={'precision': 'bf16'}).to(dev)
tr = s.RayTrainer(rend)
b = {k: v.to(dev) for k, v in s.rays.synthetic_rays(1024, seed=0).items()}
rb = {k: v.to(dev) for k, v in s.rays.synthetic_rays(4096, seed=1).items()}
for i in range(3):
    tr.step(b['rays_o'], b['rays_d'], b['times'], b['target'])
    with torch.no_grad():
        rend(rb['rays_o'], rb['rays_d'], rb['times'])
torch.cuda.synchronize()
L = _lib.lib()
buf = (ctypes.c_ulonglong * (2 * 148 * 8))()
L.snf_debug_prof.restype = ctypes.c_int
L.snf_debug_prof.argtypes = [ctypes.c_void_p]
assert L.snf_debug_prof(buf) == 0
a = np.array(buf, dtype=np.int64).reshape(2, 148, 8)
names = ['issuer total', 'issuer wait ready', 'issuer wait full', 'epi total', 'epi wait acc0', 'epi wait acc1', 'epi enc', 'epi wait stores']
for which, tag in ((0, 'inference (last launch: fine pass, 4096 rays)'), (1, 'training (last launch: fine pass, 1024 rays)')):
    print(tag)
    lead = a[which, 0::2]
    for i, n in enumerate(names[:8]):
        v = lead[:, i]
        print(f'  {n:20s} mean {v.mean():12.0f}  min {v.min():12d}  max {v.max():12d}')
tb = (ctypes.c_longlong * (4 * 512))()
L.snf_debug_trace.restype = ctypes.c_int
L.snf_debug_trace.argtypes = [ctypes.c_void_p]
assert L.snf_debug_trace(tb) == 0
t = np.array(tb, dtype=np.int64).reshape(4, 512)
t0 = t[1, 0]
t0 = t[1, 0]
print('issuer / producer trace (cycles since the first full-wait of CTA 0, second tile)')
print('blk: producer saw empty | issuer starts full-wait | done')
for i in range(0, 116):
    print(f'{i:4d} {t[0, i] - t0:9d} {t[1, i] - t0:9d} {t[2, i] - t0:9d}  wait {t[2, i] - t[1, i]:6d}')
print('epilogue warp 0 stamps: per layer [acc0 done, acc1 done, ready0 arrive, ready1..4 arrive] (enc ready first)')
e = t[3]
n = int((e != 0).sum())
print(' '.join(str(int(x - t0)) for x in e[:n]))
