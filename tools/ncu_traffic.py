"""Per-kernel DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of one `ncu --set full` capture ->
profiles/<name>.json; bench.py reports the sum over the field-network kernels of one training step as roofline.traffic.

    python tools/ncu_traffic.py gpurun_out/prof_mlp_xxx.ncu-rep profiles/r01_ncu_traffic.json
"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
col = lambda name: next(i for i, x in enumerate(h) if x == name)
kn, rd, wr, du = col('Kernel Name'), col('dram__bytes_read.sum'), col('dram__bytes_write.sum'), col('gpu__time_duration.sum')
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
tscale = {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0}
ks = []
for r in rows[2:]:
    ks.append({'kernel': r[kn].split('(')[0], 'dram_read_bytes': float(r[rd]) * scale[units[rd]],
               'dram_write_bytes': float(r[wr]) * scale[units[wr]], 'duration_s': float(r[du]) * tscale[units[du]]})
is_train_fwd = lambda n: 'mlp_fwd_bf16_kernel' in n and not ('<0>' in n or '(bool)0' in n or 'false' in n)
train = [k for k in ks if is_train_fwd(k['kernel']) or 'dgrad' in k['kernel'] or 'wgrad' in k['kernel']]
doc = {'source': rep, 'kernels': ks,
       'train_step_mlp_traffic_bytes': sum(k['dram_read_bytes'] + k['dram_write_bytes'] for k in train),
       'train_step_mlp_kernels': len(train)}
json.dump(doc, open(out, 'w'), indent=1)
print(json.dumps({k: v for k, v in doc.items() if k != 'kernels'}))
