#!/bin/bash
# weak-scaling series on one multi-GPU box: tools/gpu_scaling.sh "8 1" -> gpurun_out/scale_N.json (one bench line each)
mkdir -p gpurun_out
for n in ${1:-1 2 4 8}; do
  if [ "$n" = 1 ]; then
    timeout 500 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/scale_$n.err | tail -1 > gpurun_out/scale_$n.json
  else
    timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/scale_$n.err | tail -1 > gpurun_out/scale_$n.json
  fi
  python -c "
import json; d = json.load(open('gpurun_out/scale_$n.json')); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'], 4), round(d['e2e']['value']), d['clocks'])"
done
