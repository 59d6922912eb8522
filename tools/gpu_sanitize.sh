#!/bin/bash
# compute-sanitizer over the tensor-core kernels and the ragged / empty shapes (one --tool per gpurun call, see
# /opt/skills/guides/B200_PROFILING.md).  usage: gpurun --timeout 1500 -- bash tools/gpu_sanitize.sh memcheck|racecheck
tool=${1:-memcheck}
out=gpurun_out; mkdir -p $out
sel="field_network_forward or field_network_ragged_and_empty or emission_training_gradients_bf16 or emission_training_gradients_x3 or ray_kernels_ragged or ragged_batch_and_padding or hierarchical_sampler_perturb or spherical_sampler"
# the same selection without the tool first: it must pass on its own
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$sel" > $out/san_plain.log 2>&1 || { tail -20 $out/san_plain.log; exit 1; }
tail -2 $out/san_plain.log
extra=""
[ "$tool" = racecheck ] && extra="--racecheck-report all"
timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $tool $extra --log-file $out/sanitizer_$tool.log --print-limit 50 \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$sel" > $out/san_$tool.pytest.log 2>&1
echo "sanitizer($tool) rc=$?"
tail -3 $out/san_$tool.pytest.log
grep -c "========= " $out/sanitizer_$tool.log; tail -15 $out/sanitizer_$tool.log
