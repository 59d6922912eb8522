#!/bin/bash
# ncu --set full of the sampling / compositing kernels at N = 2^20 rays, summarised on the box.
# usage: tools/gpu_prof_hbm.sh <tag> <bench --only filter> <ncu kernel regex>
tag=${1:-x}; only=${2:-dt}; rx=${3:-composite_dt}
mkdir -p gpurun_out
timeout 120 python tools/bench_hbm_kernels.py --only $only --iters 3 > gpurun_out/hbm_plain_$tag.log 2>&1 || { tail gpurun_out/hbm_plain_$tag.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 2 -o /tmp/prof_hbm_$tag -f \
    python tools/bench_hbm_kernels.py --only $only --iters 3 > gpurun_out/ncu_hbm_$tag.log 2>&1
python tools/ncu_summary.py /tmp/prof_hbm_$tag.ncu-rep 25 > gpurun_out/ncu_hbm_summary_$tag.txt 2>&1
ncu -i /tmp/prof_hbm_$tag.ncu-rep --page raw --csv > /tmp/prof_hbm_$tag.csv 2>/dev/null
python - /tmp/prof_hbm_$tag.csv <<'P' >> gpurun_out/ncu_hbm_summary_$tag.txt
import csv, sys, io
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ['smsp__inst_executed.sum', 'smsp__inst_executed.avg', 'sm__inst_executed.avg.per_cycle_elapsed', 'sm__throughput', 'smsp__inst_issued', 'sm__inst_executed_pipe', 'sm__inst_executed_pipe_xu', 'sm__inst_executed_pipe_fp64', 'sm__inst_executed_pipe_lsu', 'sm__inst_executed_pipe_alu',
        'sm__inst_executed_pipe_fma', 'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct_of_peak', 'smsp__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared', 'sm__pipe_fp64_cycles_active', 'sm__pipe_xu_cycles_active', 'smsp__inst_executed_pipe_uniform']
print('== counters')
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')][:60])
    for i, h in enumerate(hdr):
        if any(h.startswith(w) for w in want):
            print(f'   {h} = {r[i]} {rows[1][i]}')
P
grep -A60 '== counters' gpurun_out/ncu_hbm_summary_$tag.txt | cut -c1-160; head -64 gpurun_out/ncu_hbm_summary_$tag.txt | cut -c1-170
