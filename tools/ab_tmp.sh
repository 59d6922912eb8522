timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do for f in "" "--serial-backward"; do
timeout 300 python bench.py --steps 30 --warmup 5 $f 2>&1 | tail -1 > gpurun_out/x.json; python -c "
import json;d=json.load(open('gpurun_out/x.json'));print('$f', round(d['value']),d['ms_per_step'],[round(k['ms_per_step'],3) for k in d['roofline']['kernels']], d['e2e']['value'], d['gpu_launches'], d['clocks']['sm_mhz'])"; done; done
