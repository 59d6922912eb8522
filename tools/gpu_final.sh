#!/bin/bash
# end-of-round measurement set (one GPU): parity suite, smoke, both bench arms, launch list, ncu --set full of the MLP kernels, HBM kernels
tag=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_$tag.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference_$tag.json 2> gpurun_out/bench_reference_$tag.err
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
timeout 600 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32_$tag.json 2> gpurun_out/bench_fp32_$tag.err
timeout 400 python tools/bench_hbm_kernels.py --json gpurun_out/hbm_kernels_$tag.json > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_ -s 8 -c 8 -f -o gpurun_out/prof_mlp_$tag python tools/prof_step.py 2 both > gpurun_out/ncu_full_$tag.log 2>&1
cat gpurun_out/pytest_gpu_$tag.log; tail -2 gpurun_out/smoke_$tag.log; cat gpurun_out/bench_reference_$tag.json | cut -c 1-400; cat gpurun_out/bench_$tag.json
