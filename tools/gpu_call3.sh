#!/bin/bash
mkdir -p gpurun_out
python tools/prof_step.py 3 both > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/prof_step.py 3 both > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_ -s 8 -c 6 -o gpurun_out/prof_mlp_r1 python tools/prof_step.py 3 both > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/plain.log; tail -3 gpurun_out/ncu1.log; tail -3 gpurun_out/ncu2.log; ls -la gpurun_out
