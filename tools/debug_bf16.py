import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sunerf_b200 as s
from sunerf_b200 import ops
torch.manual_seed(0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2560
net = s.NeRF(precision='bf16').cuda()
ps = net.linear_params(); W, B = ps[0::2], ps[1::2]
x = torch.randn(M, 4).cuda()
pk = net._packed_ptr(W, B); torch.cuda.synchronize(); print('pack ok', flush=True)
out, ws = ops.mlp_forward(x, W, B, mode='bf16', train=False, packed_ptr=pk); torch.cuda.synchronize(); print('fwd infer ok', out.abs().mean().item(), flush=True)
out2, ws = ops.mlp_forward(x, W, B, mode='bf16', train=True, packed_ptr=pk); torch.cuda.synchronize(); print('fwd train ok', (out - out2).abs().max().item(), flush=True)
g = torch.randn(M, 2).cuda()
gW = [torch.zeros_like(w) for w in W]; gB = [torch.zeros(w.shape[0]).cuda() for w in W]
ops.mlp_backward(x, W, g, ws, gW, gB, packed_ptr=pk); torch.cuda.synchronize(); print('bwd ok', [float(t.norm()) for t in gW][:3], flush=True)
