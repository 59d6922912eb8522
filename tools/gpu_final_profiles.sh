#!/bin/bash
# Final evidence of a round: ncu launch list of the bench command and a full-set summary of the HBM-class kernels.
tag=${1:-r02}
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_short.json 2> gpurun_out/${tag}_bench_short.err || { tail -5 gpurun_out/${tag}_bench_short.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cuda-graph > gpurun_out/${tag}_ncu_launch.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/${tag}_launches.csv
bash tools/gpu_prof_hbm.sh ${tag}_all "" "stratified_kernel|hier_kernel|composite_" 0 400 8 > /dev/null 2>&1
ls -la gpurun_out/ncu_hbm_summary_${tag}_all.txt
