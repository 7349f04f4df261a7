#!/bin/bash
# round-2 GPU visit w: BASELINE.json configs[4] on one GPU — cfg5: 3M Gaussians, 3840x2160, camera batch 32 per iteration
TAG=r02w; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 280 python bench.py --config cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/${TAG}_bench_cfg5.json 2> $OUT/${TAG}_bench_cfg5.err
echo "rc=$? t=${SECONDS}s"; head -c 900 $OUT/${TAG}_bench_cfg5.json; echo; grep -v "^\s*$" $OUT/${TAG}_bench_cfg5.err | tail -4 | cut -c1-300
nvidia-smi --query-gpu=memory.used,memory.total --format=csv,noheader
echo "elapsed ${SECONDS}s"
