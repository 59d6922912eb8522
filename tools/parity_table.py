"""Markdown table of the measured parity maxima from tests/test_gpu_config_scale.py's report (gpurun_out/config_scale_report.json)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'config_scale_report.json')))
f = lambda v: f'{v:.1e}'
names = {'fp32': 'fp32 (SIMT)', 'x3': 'x3 (split fp16, tcgen05)', 'bf16': '16-bit (tcgen05)'}
print('| workload | mode | coarse image | fine image | loss | worst gradient tensor | median | grad norm | Adam kernel vs torch | params after step |')
print('|---|---|---|---|---|---|---|---|---|---|')
for kind, label in (('emission', 'emission train step, 1024 rays'), ('dt', 'DT_2012_11 train step, 3072 rays, C=7')):
    for p in ('fp32', 'x3', 'bf16'):
        r = d.get(f'train/{kind}/{p}')
        if r:
            print(f"| {label} | {names[p]} | {f(r['coarse_image_max_rel'])} | {f(r['fine_image_max_rel'])} | {f(r['loss_rel'])} | "
                  f"{f(r['grad_worst_rel_l2'])} ({r['grad_worst_tensor']}) | {f(r['grad_median_rel_l2'])} | {f(r['grad_norm_rel'])} | "
                  f"{f(r['adam_kernel_vs_torch_on_same_grads_worst_rel_l2'])} | {f(r['param_after_step_worst_rel_l2'])} |")
print()
print('| render batch, 4096 rays | mode | coarse image | fine image (reference sum) | pixels outside gate | fine image (exact-sum oracle) | moved depths, max |')
print('|---|---|---|---|---|---|---|')
for key, label in (('emission/nerf', 'emission'), ('dt/nerf', 'render_mhd, NeRF_DT, C=6'), ('dt/simple_star', 'render_mhd, SimpleStar, C=6')):
    for p in ('fp32', 'x3', 'bf16'):
        r = d.get(f'render4096/{key}/{p}')
        if r:
            print(f"| {label} | {names[p]} | {f(r['coarse_image_max_rel'])} | {f(r['fine_image_max_rel'])} | {r['pixels_outside_gate_vs_reference_sum']} | "
                  f"{f(r['fine_image_max_rel_exact_sum_oracle'])} | {f(r['new_z_max_abs'])} |")
