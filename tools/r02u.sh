#!/bin/bash
# round-2 GPU visit u: ncu launch list of graph replays at the final code state (cfg4), hull on one GPU for the N-GPU comparison
TAG=r02u; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
   --log-file $OUT/${TAG}_launches_graph_cfg4.csv env FSB_PROFILE=cfg4 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/${TAG}_ncu_launch.log 2>&1
tail -2 $OUT/${TAG}_ncu_launch.log | cut -c1-200; echo "t=${SECONDS}s"
FSB_NO_CPU=1 timeout 200 python tools/hull_bench.py 512 5 > $OUT/${TAG}_hull_n1.json 2> $OUT/${TAG}_hull_n1.err; cut -c1-300 $OUT/${TAG}_hull_n1.json; tail -2 $OUT/${TAG}_hull_n1.err | cut -c1-200
echo "elapsed ${SECONDS}s"
