"""Developer study (CPU, oracle arithmetic): which rounding of the bf16 tensor-core training path costs how much
per-parameter GRADIENT accuracy (north_star gate: 1e-3 relative per parameter tensor)?

Emulates the field network's forward / dgrad / wgrad with a configurable rounding at every place the kernels round:
  w    : weights as MMA operands          ('bf16' | 'split' = hi+lo bf16 pair (3-MMA compensated) | 'fp32')
  h    : saved hidden activations h=sin   ('bf16' | 'split' | 'fp32')   (operand of the next layer and of the wgrad)
  cos  : cos(pre) used by the dgrad chain ('int8' | 'bf16' | 'u16phase' | 'fp32')
  dpre : dL/dpre as operand of dgrad/wgrad ('bf16' | 'split' | 'fp32')
and runs the oracle's emission train step around it.  Prints max over the 36 parameter tensors of
||g - g_ref|| / ||g_ref|| plus the intensity error."""
import os, sys, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import sunerf_oracle as orc

torch.set_num_threads(os.cpu_count() or 1)
F = torch.nn.functional


def bf(t):
    return t.to(torch.bfloat16).float()


def rnd(t, kind):
    if kind == 'fp32':
        return t
    if kind == 'bf16':
        return bf(t)
    if kind == 'fp16':
        return t.to(torch.float16).float()
    if kind == 'split16':                   # hi + lo fp16 pair (lo may be subnormal: absolute floor 2^-25)
        hi = t.to(torch.float16).float()
        return hi + (t - hi).to(torch.float16).float()
    if kind == 'split':                     # hi + lo bf16 pair: 16 mantissa bits
        hi = bf(t)
        return hi + bf(t - hi)
    raise ValueError(kind)


def cos_of(pre, kind):
    c = torch.cos(pre)
    if kind == 'fp32':
        return c
    if kind == 'int8':
        return torch.round(c * 127.0) / 127.0
    if kind == 'bf16':
        return bf(c)
    if kind == 'fp16':
        return c.to(torch.float16).float()
    if kind == 'i8half':                    # 7-bit magnitude of sqrt(1-|c|) (half-angle code) + sign bit
        tq = torch.round(127.0 * torch.sqrt(1 - c.abs())) / 127.0
        return torch.sign(c) * (1 - tq * tq)
    if kind == 'i8half_k':                  # the kernels' arithmetic: encode with T = 127.014, decode in fp16 pairs
        T = 127.0140556
        q = torch.round(T * torch.sqrt(1 - c.abs())).clamp(max=127)
        h = lambda t_: t_.to(torch.float16).float()
        q2 = h(q * q)
        mag = h(q2 * h(torch.tensor(-1.0 / (T * T))) + 1.0)
        return torch.sign(c) * mag
    if kind == 'i8angle':                   # 8-bit folded phase acos(c) in [0, pi]
        ph = torch.round(torch.acos(c.clamp(-1, 1)) * (255.0 / torch.pi)) * (torch.pi / 255.0)
        return torch.cos(ph)
    if kind == 'u16phase':                  # phase quantised to 2*pi/65536
        q = torch.round(pre * (65536.0 / (2 * torch.pi)))
        return torch.cos(q * (2 * torch.pi / 65536.0))
    raise ValueError(kind)


class EmuMLP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cfg, *params):
        ws, bs = params[0::2], params[1::2]
        n = len(ws)
        h = rnd(orc.positional_encoding(x), 'bf16' if cfg['h'] == 'u16phase' else cfg['h'])
        hs, coss = [h], []
        ctx_full = []
        for i in range(n - 1):
            pre = F.linear(h, rnd(ws[i], cfg['w'])) + bs[i]
            coss.append(cos_of(pre, cfg['cos']))
            h_full = torch.sin(pre)
            ctx_full.append(h_full)
            if cfg['h'] == 'u16phase':
                q = torch.round(pre * (65536.0 / (2 * torch.pi)))
                h = torch.sin(q * (2 * torch.pi / 65536.0))
            else:
                h = rnd(h_full, cfg['h'])
            hs.append(h)
        out = F.linear(h_full, ws[-1], bs[-1])          # output layer: fp32 registers on the unrounded h_7
        if 'h_bwd' in cfg:                      # the backward reads another rounding of the saved activations / weights
            hs = [rnd(orc.positional_encoding(x), cfg['h_bwd'])] + [rnd(t_, cfg['h_bwd']) for t_ in ctx_full]
        ctx.cfg, ctx.hs, ctx.coss, ctx.ws, ctx.h7 = cfg, hs, coss, ws, h_full
        return out

    @staticmethod
    def backward(ctx, g):
        cfg, hs, coss, ws = ctx.cfg, ctx.hs, ctx.coss, ctx.ws
        n = len(ws)
        grads = [None] * (2 * n)
        grads[2 * (n - 1)] = g.t() @ hs[-1]
        grads[2 * (n - 1) + 1] = g.sum(0)
        dh = g @ ws[-1]
        S = 1.0
        if cfg['dpre'] in ('fp16s', 'fp16s2'):  # fp16 with one power-of-two scale per call (GradScale of the kernels)
            import math
            bound = g.abs().max().item() * (ws[-1][0].abs() + ws[-1][1].abs()).max().item()
            S = 2.0 ** (-math.frexp(bound)[1]) if bound > 0 else 1.0
        for i in range(n - 2, -1, -1):
            if cfg['dpre'] == 'fp16s2':         # the kernels' double rounding: fp16(fp16(S dh) * cos)
                dpre = ((dh * S).to(torch.float16).float() * coss[i]).to(torch.float16).float() / S
            elif cfg['dpre'] == 'fp16s':
                dpre = (dh * coss[i] * S).to(torch.float16).float() / S
            else:
                dpre = rnd(dh * coss[i], cfg['dpre'])
            grads[2 * i] = dpre.t() @ hs[i]
            grads[2 * i + 1] = dpre.sum(0)
            if i > 0:
                dh = dpre @ rnd(ws[i], cfg.get('w_bwd', cfg['w']))
        return (None, None) + tuple(grads)


KIND = 'emission'
TARGET_MODE = 'centred'


def run(cfg, b, seeds=(1, 2), scale=1.0):
    if KIND == 'dt':
        return run_dt(cfg, b, seeds)
    pc, pf = orc.FieldParams.init(seeds[0]), orc.FieldParams.init(seeds[1])
    for p in (pc, pf):
        for i in range(1, 8):
            p.weights[i] = p.weights[i] * scale
    pc.requires_grad_(); pf.requires_grad_()
    orig = orc.field_mlp
    if cfg is not None:
        orc.field_mlp = lambda x, p, *a, **k: EmuMLP.apply(x, cfg, *p.tensors())
    try:
        rc = orc.RenderConfig(kind='emission')
        t_rand = torch.rand(b['rays_o'].shape[0], 64, generator=torch.Generator().manual_seed(3))
        out = orc.render(rc, pc, pf, b['rays_o'], b['rays_d'], b['times'], None, t_rand)
        loss = orc.training_loss(out, b['target'], 'emission')['loss']
        loss.backward()
    finally:
        orc.field_mlp = orig
    return out['fine_image'].detach(), [t.grad.clone() for t in pc.tensors() + pf.tensors()]


_dt_target = {}


def run_dt(cfg, b, seeds):
    import numpy as np
    a = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests', 'golden', 'aia_response.npz'))
    tx, ty = torch.from_numpy(a['logT'].copy()), torch.from_numpy(a['table'].copy())
    pc, pf = orc.FieldParams.init(seeds[0], dt=True), orc.FieldParams.init(seeds[1], dt=True)
    pc.log_abs = torch.tensor([2e-6 * (i + 1) for i in range(7)]); pf.log_abs = torch.tensor([3e-6 * (i + 1) for i in range(7)])
    pc.requires_grad_(); pf.requires_grad_()
    n = b['rays_o'].shape[0]
    wl = torch.tensor([94., 131., 171., 193., 211., 304., 335.]).repeat(n, 1)
    wl[n // 2:] = torch.tensor([0., 0., 171., 193., 211., 304., 0.])
    rc = orc.RenderConfig(kind='dt', pixel_intensity_factor=1e17, table_x=tx, table_y=ty)
    t_rand = torch.rand(n, 64, generator=torch.Generator().manual_seed(3))
    orig = orc.field_mlp
    if cfg is not None:
        def emu(x, p, o0=0.0, o1=0.0):
            out = EmuMLP.apply(x, cfg, *p.tensors()[:18])
            return torch.stack([out[:, 0] + o0, out[:, 1] + o1], -1)
        orc.field_mlp = emu
    try:
        out = orc.render(rc, pc, pf, b['rays_o'], b['rays_d'], b['times'], wl, t_rand)
        if 't' not in _dt_target:
            u = torch.rand(n, 7, generator=torch.Generator().manual_seed(9))
            f = (0.5 + u) if TARGET_MODE == 'centred' else (0.25 + 0.5 * u)
            _dt_target['t'] = (out['fine_image'].detach() * f) * (wl > 0)
        loss = orc.training_loss(out, _dt_target['t'], 'dt')['loss']
        loss.backward()
    finally:
        orc.field_mlp = orig
    return out['fine_image'].detach(), [t.grad.clone() for t in pc.tensors() + pf.tensors()]


def main():
    global KIND, TARGET_MODE
    if len(sys.argv) > 2:
        KIND = sys.argv[2]
    if len(sys.argv) > 3:
        TARGET_MODE = sys.argv[3]
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    b = orc.synthetic_rays(n, seed=5)
    img_ref, g_ref = run(None, b)
    base = dict(w='bf16', h='bf16', cos='int8', dpre='bf16')
    f16 = dict(w='fp16', h='fp16', cos='fp16', dpre='bf16')
    x3 = dict(w='split16', h='split16', cos='i8half', dpre='fp16s', w_bwd='split16', h_bwd='fp16')
    variants = [('x3 ideal decode', x3),
                ('x3 fp16 decode', dict(x3, cos='i8half_k')),
                ('x3 fp16 decode + double rounding', dict(x3, cos='i8half_k', dpre='fp16s2')),
                ('x3 ideal decode + double rounding', dict(x3, dpre='fp16s2')),
                ('x3 exact cos', dict(x3, cos='fp32')),
                ('x3 exact cos, h_bwd fp32', dict(x3, cos='fp32', h_bwd='fp32')),
                ]
    if os.environ.get('STUDY_ONLY'):
        variants = [v for v in variants if v[0].startswith(tuple(os.environ['STUDY_ONLY'].split(',')))]
    for name, cfg in variants:
        img, g = run(cfg, b)
        ie = ((img - img_ref).abs() / (img_ref.abs() + 1e-30)).max().item()
        errs = [((a - r).norm() / r.norm()).item() for a, r in zip(g, g_ref)]
        werr = max(errs[0::2]); berr = max(errs[1::2])
        if os.environ.get('STUDY_VERBOSE'):
            print('   per tensor:', ' '.join(f'{e:.1e}' for e in errs))
        print(f'{name:34s} intensity {ie:.2e}  grad W max {werr:.2e}  grad b max {berr:.2e}  median {sorted(errs)[len(errs)//2]:.2e}', flush=True)


if __name__ == '__main__':
    main()
