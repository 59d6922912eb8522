#!/bin/bash
# Developer aid: A/B of the multi-GPU step schedule on N GPUs (usage: gpurun --gpus N -- bash tools/scale_ab.sh N)
N=${1:-2}
run() { env "$@" SNF_BENCH_WATCHDOG=60 timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
        bench.py --gpus $N --steps 40 --warmup 5 --quick 2>/dev/null | grep ms_per_step | sed "s/^/$* /"; }
timeout 90 python bench.py --steps 40 --warmup 5 --quick 2>/dev/null | grep ms_per_step
for rep in 1; do
run SNF_EARLY_REDUCE=0 SNF_RESERVE_SMS=0
run SNF_EARLY_REDUCE=1 SNF_RESERVE_SMS=0
run SNF_EARLY_REDUCE=1 SNF_RESERVE_SMS=4
run SNF_EARLY_REDUCE=0 SNF_RESERVE_SMS=4
done
