#!/bin/bash
# 2-GPU visit: N=2 bench (graph, then eager) under torchrun exactly as the driver launches it.
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
SECONDS=0
timeout 400 $TR bench.py --gpus 2 --steps 50 --warmup 5 > $OUT/${TAG}_bench_graph_n2.json 2> $OUT/${TAG}_bench_graph_n2.err; echo "rc $? t=${SECONDS}s"; tail -c 400 $OUT/${TAG}_bench_graph_n2.json; tail -5 $OUT/${TAG}_bench_graph_n2.err
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 --mode eager > $OUT/${TAG}_bench_eager_n2.json 2> $OUT/${TAG}_bench_eager_n2.err; echo "rc $? t=${SECONDS}s"; tail -c 300 $OUT/${TAG}_bench_eager_n2.json
ls -la $OUT | tail -6
