"""Developer study (CPU, oracle arithmetic): what would fp8 (e4m3) hidden activations cost in accuracy?

Emulates the tensor-core path of the field network - bf16 weights, fp32 accumulation - with the hidden activations
h = sin(pre) rounded to bf16 (what the kernels do today) or to e4m3 in the layers listed, and renders the golden emission
rays with both networks.  Reports the relative intensity error against the fp32 render (gate of the bf16-MLP mode: 1e-2)
for the default initialisation and for scaled hidden weights ("limit of the bf16 mode")."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import sunerf_oracle as orc

torch.set_num_threads(os.cpu_count() or 1)


def q(t, kind):
    if kind == 'fp32':
        return t
    if kind == 'bf16':
        return t.to(torch.bfloat16).float()
    return t.to(torch.float8_e4m3fn).float()


def qw(w, kind):
    """weights: bf16, or e4m3 with one scale per output feature (folded into the epilogue as pre = s_n * acc + b_n)"""
    if kind != 'e4m3w':
        return q(w, 'bf16')
    s = w.abs().amax(dim=1, keepdim=True).clamp_min(1e-30) / 448.0
    return (w / s).to(torch.float8_e4m3fn).float() * s


def mlp(x, p, kinds):
    """kinds[i]: precision of the OUTPUT activations of layer i; a trailing 'w' (e4m3w) means the layer that CONSUMES them
    also takes its weights in e4m3 (tcgen05 kind::f8f6f4 needs both operands in 8 bits)."""
    h = q(orc.positional_encoding(x), 'bf16')
    n = len(p.weights)
    for i in range(n - 1):
        wk = 'e4m3w' if (i > 0 and kinds[i - 1] == 'e4m3w') else 'bf16'
        h = q(torch.sin(torch.nn.functional.linear(h, qw(p.weights[i], wk), p.biases[i])), kinds[i].rstrip('w'))
    return torch.nn.functional.linear(h, p.weights[-1], p.biases[-1])


def render(cfg, pc, pf, b, kinds):
    orig = orc.field_mlp
    orc.field_mlp = (lambda x, p, *a, **k: mlp(x, p, kinds)) if kinds is not None else orig
    try:
        with torch.no_grad():
            return orc.render(cfg, pc, pf, b['rays_o'], b['rays_d'], b['times'], None, None)
    finally:
        orc.field_mlp = orig


def main():
    b = orc.synthetic_rays(512, seed=5)
    cfg = orc.RenderConfig(kind='emission')
    for scale in (1.0, 1.5, 2.0):
        pc, pf = orc.FieldParams.init(1), orc.FieldParams.init(2)
        for p in (pc, pf):
            for i in range(1, 8):
                p.weights[i] = p.weights[i] * scale
        ref = render(cfg, pc, pf, b, None)['fine_image']
        rows = []
        for name, kinds in (('bf16 activations (today)', ['bf16'] * 8),
                            ('e4m3 in layers 1-7', ['bf16'] + ['e4m3'] * 7),
                            ('e4m3 in layers 4-7', ['bf16'] * 4 + ['e4m3'] * 4),
                            ('e4m3 in layer 7 only', ['bf16'] * 7 + ['e4m3']),
                            ('e4m3 activations AND weights, layers 1-7', ['e4m3w'] * 7 + ['bf16']),
                            ('e4m3 activations AND weights, layers 4-7', ['bf16'] * 3 + ['e4m3w'] * 4 + ['bf16'])):
            img = render(cfg, pc, pf, b, kinds)['fine_image']
            rows.append((name, ((img - ref).abs() / ref.abs().clamp_min(1e-30)).max().item()))
        print(f'hidden weights x{scale}: ' + '; '.join(f'{n}: {e:.2e}' for n, e in rows), flush=True)


if __name__ == '__main__':
    main()
