#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "bf16 or ragged or full_size" 2>&1 | tail -40 > gpurun_out/pytest_bf16.log
timeout 600 python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1
tail -40 gpurun_out/pytest_bf16.log; tail -3 gpurun_out/bench_bf16.log
