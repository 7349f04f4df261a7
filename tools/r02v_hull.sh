#!/bin/bash
# round-2 GPU visit v: VisualHull 512^3 z-slab sharded over N GPUs (BASELINE.json configs[2])
N=${1:-4}; TAG=r02v; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/hull_bench.py 512 5 > $OUT/${TAG}_hull_n${N}.json 2> $OUT/${TAG}_hull_n${N}.err
echo "rc=$?"; cut -c1-400 $OUT/${TAG}_hull_n${N}.json; grep -v "^\s*$" $OUT/${TAG}_hull_n${N}.err | tail -2 | cut -c1-200
echo "elapsed ${SECONDS}s"
