#!/bin/bash
# 2-GPU visit: N=2 bench (graph mode) under torchrun exactly as the driver launches it.
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
SECONDS=0
timeout 300 $TR bench.py --gpus 2 --steps 50 --warmup 5 > $OUT/${TAG}_bench_graph_n2.json 2> $OUT/${TAG}_bench_graph_n2.err; echo "rc $? t=${SECONDS}s"; head -c 300 $OUT/${TAG}_bench_graph_n2.json; echo; tail -3 $OUT/${TAG}_bench_graph_n2.err | cut -c1-300
