#!/bin/bash
# Developer aid: N-GPU step time against the 1-GPU step on the same box, default schedule vs no SM reserve / late reduce
N=${1:-8}
run() { env "$@" SNF_BENCH_WATCHDOG=80 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
        bench.py --gpus $N --steps 40 --warmup 5 --quick 2>/dev/null | grep ms_per_step | sed "s/^/$* /"; }
timeout 90 python bench.py --steps 40 --warmup 5 --quick 2>/dev/null | grep ms_per_step
run SNF_EARLY_REDUCE=1 SNF_RESERVE_SMS=4
run SNF_EARLY_REDUCE=0 SNF_RESERVE_SMS=0
