#!/bin/bash
# round-2 visit y: bench.py at 2 GPUs as the driver launches it, final code state (e2e leg with its own warm-up)
N=2; TAG=r02y; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 60 --warmup 5 --no-cpu-baseline --no-secondary > $OUT/${TAG}_bench_n${N}.json 2> $OUT/${TAG}_bench_n${N}.err
echo "rc=$? t=${SECONDS}s"; head -c 2500 $OUT/${TAG}_bench_n${N}.json | grep -o '"value": [0-9.]*\|"h2d_bytes_per_step": [0-9]*\|"ms_per_step": [0-9.]*' | head -6; grep -v "^\s*$" $OUT/${TAG}_bench_n${N}.err | tail -2 | cut -c1-300
echo "elapsed ${SECONDS}s"
