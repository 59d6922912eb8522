#!/bin/bash
mkdir -p gpurun_out
python tools/prof_step.py 2 both > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp_ -s 6 -c 6 -o gpurun_out/prof_mlp_r1b python tools/prof_step.py 2 both > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/plain.log; tail -3 gpurun_out/ncu2.log
