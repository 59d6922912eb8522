"""Developer diagnostic: per-tensor gradient error of the split-precision mode with the dgrad chain's W^T operand as one
fp16 (16-bit backward) and as the (hi, lo) pair, against the oracle, on the smoke batch."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sunerf_b200 as s
from sunerf_b200 import ops
from oracle import sunerf_oracle as orc

dev = torch.device('cuda', 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rays = s.rays.synthetic_rays(N, seed=3, H=64, W=64, plate_arcsec=40.0)
t_rand = torch.rand(N, 64, generator=torch.Generator().manual_seed(1))


def oracle(rend):
    def params(m):
        ws = [m.in_layer[1].weight] + [l.weight for l in m.layers] + [m.out_layer.weight]
        bs = [m.in_layer[1].bias] + [l.bias for l in m.layers] + [m.out_layer.bias]
        return orc.FieldParams([w.detach().cpu().clone() for w in ws], [b.detach().cpu().clone() for b in bs])
    pc, pf = params(rend.coarse_model).requires_grad_(), params(rend.fine_model).requires_grad_()
    ref = orc.render(orc.RenderConfig(kind='emission'), pc, pf, rays['rays_o'], rays['rays_d'], rays['times'], None, t_rand)
    orc.training_loss(ref, rays['target'], 'emission')['loss'].backward()
    return pc, pf


_bwd = ops.mlp_backward
for label, precision, force16 in (('16-bit mode', 'bf16', False), ('x3 fwd + 16-bit bwd (W^T single)', 'x3', True), ('x3 fwd + split W^T', 'x3', False)):
    torch.manual_seed(11)
    rend = s.EmissionRadiativeTransfer(Rs_per_ds=1, model_config={'precision': precision}).to(dev)
    pc, pf = oracle(rend)
    if force16:
        def patched(x, weights, grad_out, ws, gw, gb, packed_ptr=None):
            ws.mode = 'bf16'
            return _bwd(x, weights, grad_out, ws, gw, gb, packed_ptr=packed_ptr)
        import sunerf_b200.trainer as T
        ops.mlp_backward = patched
    tr = s.RayTrainer(rend)
    tr.step(*(rays[k].to(dev) for k in ('rays_o', 'rays_d', 'times', 'target')), t_rand=t_rand.to(dev))
    ops.mlp_backward = _bwd
    errs = []
    for name, p in (('coarse_model', pc), ('fine_model', pf)):
        for got, ref_t in zip(getattr(rend, name).linear_params(), p.tensors()):
            g_ref = ref_t.grad.double()
            errs.append(float((tr.grad_view[id(got)].double().cpu() - g_ref).norm() / g_ref.norm()))
    print(f'{label:36s} worst {max(errs):.2e} median {np.median(errs):.2e} | ' + ' '.join(f'{e:.1e}' for e in errs), flush=True)
