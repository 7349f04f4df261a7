#!/bin/bash
# round-2 visit r: bench.py at 8 GPUs as the driver launches it, 8-bit target staging (default) — one run, kept short
N=${1:-8}; TAG=r02r; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 200 $TR bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > $OUT/${TAG}_bench_n${N}.json 2> $OUT/${TAG}_bench_n${N}.err
echo "rc=$? t=${SECONDS}s"; head -c 2500 $OUT/${TAG}_bench_n${N}.json | grep -o '"value": [0-9.]*\|"h2d_bytes_per_step": [0-9]*\|"ms_per_step": [0-9.]*' | head -6; grep -v "^\s*$" $OUT/${TAG}_bench_n${N}.err | tail -2 | cut -c1-300
echo "elapsed ${SECONDS}s"
