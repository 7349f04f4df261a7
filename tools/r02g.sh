#!/bin/bash
# round-2 GPU visit g (2 GPUs): synchronised refinement on NCCL / peer exchange (e2), VisualHull slab sharding on hardware (e3)
TAG=r02g; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
for MODE in peer nccl; do
  timeout 300 $TR tools/train_multi_gpu_check.py $MODE > $OUT/${TAG}_train_check_${MODE}.json 2> $OUT/${TAG}_train_check_${MODE}.err; echo "rc=$?"; cat $OUT/${TAG}_train_check_${MODE}.json; tail -3 $OUT/${TAG}_train_check_${MODE}.err | cut -c1-300
done
echo "t=${SECONDS}s"
timeout 300 python tools/hull_bench.py 512 5 > $OUT/${TAG}_hull_n1.json 2> $OUT/${TAG}_hull_n1.err; echo "rc=$?"; cat $OUT/${TAG}_hull_n1.json | cut -c1-900
timeout 300 $TR tools/hull_bench.py 512 5 > $OUT/${TAG}_hull_n2.json 2> $OUT/${TAG}_hull_n2.err; echo "rc=$?"; cat $OUT/${TAG}_hull_n2.json | cut -c1-900; tail -3 $OUT/${TAG}_hull_n2.err | cut -c1-300
echo "elapsed ${SECONDS}s"
