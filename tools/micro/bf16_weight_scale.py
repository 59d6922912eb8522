"""How the bf16-MLP mode's error grows with the weight scale (trained checkpoints have larger weights than
the default init).  Compares raw network outputs and rendered intensities, bf16 vs fp32 mode, for scaled hidden weights."""
import sys, torch
sys.path.insert(0, '.')
import sunerf_b200 as s
dev = torch.device('cuda', 0)
rays = {k: v.to(dev) for k, v in s.rays.synthetic_rays(512, seed=3).items()}
for scale in (1.0, 1.5, 2.0, 3.0, 4.0):
    torch.manual_seed(11)
    r = s.EmissionRadiativeTransfer(Rs_per_ds=1, sampling_config={'type': 'stratified', 'perturb': False}).to(dev)
    with torch.no_grad():
        for m in (r.coarse_model, r.fine_model):
            for lin in m.layers:
                lin.weight.mul_(scale)
    outs = {}
    for prec in ('fp32', 'bf16'):
        r.coarse_model.precision = r.fine_model.precision = prec
        with torch.no_grad():
            outs[prec] = r(rays['rays_o'], rays['rays_d'], rays['times'])
            q = torch.randn(4096, 4, device=dev)
            outs[prec + '_raw'] = r.fine_model(q)['inferences']
    rel = ((outs['bf16']['fine_image'] - outs['fp32']['fine_image']).abs() / outs['fp32']['fine_image'].abs()).max().item()
    raw = (outs['bf16_raw'] - outs['fp32_raw']).abs().max().item()
    print(f'hidden weights x{scale}: max |raw_bf16 - raw_fp32| = {raw:.2e}, max relative intensity error = {rel:.2e}')
