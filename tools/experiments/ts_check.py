"""Developer aid: TS inference forward against the SS one (values and time)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sunerf_b200 as s
from sunerf_b200 import ops
dev = torch.device('cuda', 0)
torch.manual_seed(3)
net = s.NeRF(precision='bf16').to(dev)
for M in (256, 1000, 4096 * 192):
    x = torch.randn(M, 4, device=dev)
    outs = []
    for v in (0, 1):
        ops.fwd_variant(v)
        with torch.no_grad():
            y = net(x)['inferences']
        torch.cuda.synchronize()
        outs.append(y.clone())
    d = (outs[0] - outs[1]).abs().max().item()
    print(f'M={M}: max |SS - TS| = {d:.3e}, max |SS| = {outs[0].abs().max().item():.3f}, finite {bool(torch.isfinite(outs[1]).all())}', flush=True)
M = 4096 * 192
x = torch.randn(M, 4, device=dev)
for v in (0, 1, 0, 1):
    ops.fwd_variant(v)
    with torch.no_grad():
        for _ in range(3):
            net(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            net(x)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'variant {v}: {ms:.3f} ms per {M} points, {M / ms / 1e3:.1f} Msamples/s', flush=True)
