#!/bin/bash
# ncu --set full capture of one steady-state iteration of the bf16 MLP kernels (train step + 4096-ray render)
# usage: tools/gpu_prof.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
python tools/prof_step.py 3 both > gpurun_out/plain_$tag.log 2>&1 || { tail -20 gpurun_out/plain_$tag.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:mlp_ -s 8 -c 8 -o gpurun_out/prof_mlp_$tag -f \
    python tools/prof_step.py 2 both > gpurun_out/ncu_$tag.log 2>&1
tail -3 gpurun_out/plain_$tag.log; tail -3 gpurun_out/ncu_$tag.log
