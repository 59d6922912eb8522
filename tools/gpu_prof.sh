#!/bin/bash
# ncu --set full capture of one steady-state iteration of the tensor-core MLP kernels (train step + 4096-ray render),
# summarised ON THE BOX (gpurun copies back at most 64 MiB): per-kernel headline metrics + hottest SASS lines
# (tools/ncu_summary.py) and DRAM traffic per kernel (tools/ncu_traffic.py).   usage: tools/gpu_prof.sh <tag> [bf16|x3]
tag=${1:-x}
prec=${2:-bf16}
mkdir -p gpurun_out
timeout 300 python tools/prof_step.py 3 both $prec > gpurun_out/plain_$tag.log 2>&1 || { tail -20 gpurun_out/plain_$tag.log; exit 1; }
skip=8; cnt=8
if [ "$prec" = x3 ]; then skip=36; cnt=36; fi      # x3 per iteration: 2 x (8 layer kernels + dgrad + wgrad) + 16 render layer kernels
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:mlp_ -s $skip -c $cnt -o /tmp/prof_mlp_$tag -f \
    python tools/prof_step.py 2 both $prec > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/plain_$tag.log; tail -2 gpurun_out/ncu_$tag.log
python tools/ncu_summary.py /tmp/prof_mlp_$tag.ncu-rep 12 > gpurun_out/ncu_summary_$tag.txt 2>&1
python tools/ncu_traffic.py /tmp/prof_mlp_$tag.ncu-rep gpurun_out/ncu_traffic_$tag.json
ls -la /tmp/prof_mlp_$tag.ncu-rep
head -40 gpurun_out/ncu_summary_$tag.txt | cut -c1-230
