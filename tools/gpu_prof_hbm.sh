#!/bin/bash
# ncu --set full of the sampling / compositing kernels at N = 2^20 rays, summarised on the box.
# usage: tools/gpu_prof_hbm.sh <tag> <bench --only filter> <ncu kernel regex>
#        [launches to skip] [launches to capture] [hot SASS lines per kernel]
tag=${1:-x}; only=${2-dt}; rx=${3:-composite_dt}; skip=${4:-4}; cnt=${5:-2}; top=${6:-25}
sel=""; [ -n "$only" ] && sel="--only $only"
mkdir -p gpurun_out
timeout 120 python tools/bench_hbm_kernels.py $sel --iters 3 > gpurun_out/hbm_plain_$tag.log 2>&1 || { tail gpurun_out/hbm_plain_$tag.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -o /tmp/prof_hbm_$tag -f \
    python tools/bench_hbm_kernels.py $sel --iters 3 > gpurun_out/ncu_hbm_$tag.log 2>&1
python tools/ncu_summary.py /tmp/prof_hbm_$tag.ncu-rep $top > gpurun_out/ncu_hbm_summary_$tag.txt 2>&1
ncu -i /tmp/prof_hbm_$tag.ncu-rep --page raw --csv > /tmp/prof_hbm_$tag.csv 2>/dev/null
python - /tmp/prof_hbm_$tag.csv <<'P' >> gpurun_out/ncu_hbm_summary_$tag.txt
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = lambda name: next((i for i, h in enumerate(hdr) if h == name), None)
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size']
print('== counters (one line per captured launch)')
seen = {}
for r in rows[2:]:
    name = r[col('Kernel Name')].split('(')[0][-48:]
    key = (name, r[col('launch__grid_size')] if col('launch__grid_size') is not None else '')
    seen[key] = seen.get(key, 0) + 1
    if seen[key] > 1:
        continue          # the first launch of each (kernel, grid) only
    print(name)
    for w in want:
        i = col(w)
        if i is not None:
            print(f'   {w} = {r[i]} {units[i]}')
P
grep -A60 '== counters' gpurun_out/ncu_hbm_summary_$tag.txt | cut -c1-160; head -64 gpurun_out/ncu_hbm_summary_$tag.txt | cut -c1-170
