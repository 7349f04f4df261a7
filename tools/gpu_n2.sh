#!/bin/bash
# 2-GPU visit: N=2 bench (graph + eager) under torchrun, reference arm, then one ncu --set full capture of raster backward (1 GPU).
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 50 --warmup 5 > $OUT/${TAG}_bench_graph_n2.json 2> $OUT/${TAG}_bench_graph_n2.err; echo "rc $?"; tail -c 400 $OUT/${TAG}_bench_graph_n2.json; tail -5 $OUT/${TAG}_bench_graph_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 50 --warmup 5 --mode eager > $OUT/${TAG}_bench_eager_n2.json 2> $OUT/${TAG}_bench_eager_n2.err; echo "rc $?"; tail -c 300 $OUT/${TAG}_bench_eager_n2.json
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'raster_bwd_kernel' -c 2 \
   -o $OUT/${TAG}_raster_bwd_full -f env FSB_PROFILE=1 python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline > $OUT/${TAG}_ncu_bwd_full.log 2>&1
ls -la $OUT | tail -8
