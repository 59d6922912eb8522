#!/bin/bash
# first GPU pass: layout probes, parity tests (fp32 then bf16 in separate processes), smoke, short bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/probe.log 2>&1
for t in ss ts mn mn2; do timeout 60 tools/bin/umma_probe $t >> gpurun_out/probe.log 2>&1; echo "exit $?" >> gpurun_out/probe.log; done
timeout 900 python -m pytest tests -m gpu -q -k "not bf16 and not full_size" 2>&1 | tail -60 > gpurun_out/pytest_fp32.log
timeout 600 python -m pytest tests -m gpu -q -k "bf16 or full_size" 2>&1 | tail -60 > gpurun_out/pytest_bf16.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --precision fp32 --steps 5 --warmup 3 > gpurun_out/bench_fp32.log 2>&1
cat gpurun_out/probe.log; tail -30 gpurun_out/pytest_fp32.log; tail -30 gpurun_out/pytest_bf16.log; tail -12 gpurun_out/smoke.log; tail -3 gpurun_out/bench_fp32.log
