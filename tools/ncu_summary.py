"""Summarise an .ncu-rep: per-kernel headline metrics + hottest SASS lines with stall reasons."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'sm__cycles_elapsed.avg.per_second']
idx = {w: next(i for i, h in enumerate(hdr) if h.startswith(w)) for w in want}
print('== kernels')
for r in rows[2:]:
    print(' | '.join(f'{r[idx[w]]}{rows[1][idx[w]] if w != "Kernel Name" else ""}' for w in want))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
insts, cur, name = [], None, None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == 'Kernel Name':
        cur = []; insts.append((r[1], cur)); continue
    if r and r[0] == 'Address':
        h = r; continue
    if cur is not None and len(r) > 5:
        cur.append(r)
si = h.index('Warp Stall Sampling (All Samples)')
st = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
for name, k in insts:
    tot = sum(int(r[si]) for r in k)
    print(f'== {name[:60]}  samples {tot}  instrs {len(k)}')
    top = sorted(range(len(k)), key=lambda i: -int(k[i][si]))[:top_n]
    for i in sorted(top):
        r = k[i]
        why = ', '.join(f'{h[j][6:]}={r[j]}' for j in st if r[j] not in ('0', '') and int(r[j]) * 10 > int(r[si]))
        print(f'  {i:5d} {100*int(r[si])/max(tot,1):5.1f}%  {r[1].strip()[:70]:70s} {why}')
