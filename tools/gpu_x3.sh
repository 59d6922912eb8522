#!/bin/bash
# Developer aid: first contact of the split-precision kernels with the GPU, smallest case first, each in its own process.
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "field_network_forward and x3" > $out/x3_a.log 2>&1; echo "x3 fwd: rc=$?"; tail -15 $out/x3_a.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "x3 or ragged_and_empty or sampler or hierarchical" > $out/x3_b.log 2>&1; echo "x3 all: rc=$?"; tail -30 $out/x3_b.log
