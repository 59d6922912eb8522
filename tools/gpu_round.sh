#!/bin/bash
# Developer aid: the GPU suite in separate processes (a trapped kernel poisons its CUDA context, not the next file's),
# logs under gpurun_out/.  Usage: gpurun --timeout 1500 -- bash tools/gpu_round.sh [tag]
tag=${1:-run}
out=gpurun_out
mkdir -p $out
rm -f $out/config_scale_report.json
python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "field_network" > $out/${tag}_t0_field.log 2>&1; echo "field: rc=$?"
python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 > $out/${tag}_t1_parity.log 2>&1; echo "parity: rc=$?"
python -m pytest tests/test_gpu_config_scale.py -m gpu -q -s --timeout 900 > $out/${tag}_t2_config.log 2>&1; echo "config: rc=$?"
python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_parity.py --deselect tests/test_gpu_config_scale.py > $out/${tag}_t3_rest.log 2>&1; echo "rest: rc=$?"
tail -5 $out/${tag}_t0_field.log; tail -25 $out/${tag}_t1_parity.log; grep -a "config-scale\|passed\|failed\|Error\|assert" $out/${tag}_t2_config.log | cut -c1-400 | tail -40; tail -8 $out/${tag}_t3_rest.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke: rc=$?"; tail -6 $out/${tag}_smoke.log
if [ -n "$BENCH" ]; then
  python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench: rc=$?"; tail -c 3000 $out/${tag}_bench.json; tail -5 $out/${tag}_bench.err
fi
