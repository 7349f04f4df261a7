#!/bin/bash
# round-2 GPU visit e (2 GPUs): reference dn_model.py harness; N=2 baseline (two graphs + eager NCCL) and the NCCL
# all-reduce captured inside the one graph (guarded by timeouts: it hung in round 1)
TAG=r02e; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 600 python -m pytest tests/test_reference_dn_model.py tests/test_gpu_raster_dn.py -m gpu -q > $OUT/${TAG}_pytest_ref.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_ref.log
tail -25 $OUT/${TAG}_pytest_ref.log; echo "t=${SECONDS}s"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --steps 50 --warmup 5 > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err; echo "rc=$?"; head -c 500 $OUT/${TAG}_bench_n2.json; echo; tail -3 $OUT/${TAG}_bench_n2.err | cut -c1-300; echo "t=${SECONDS}s"
FSB_CAPTURE_NCCL=1 timeout 300 $TR bench.py --gpus 2 --steps 50 --warmup 5 > $OUT/${TAG}_bench_n2_captured.json 2> $OUT/${TAG}_bench_n2_captured.err; echo "rc=$?"; head -c 500 $OUT/${TAG}_bench_n2_captured.json; echo; tail -3 $OUT/${TAG}_bench_n2_captured.err | cut -c1-300
echo "elapsed ${SECONDS}s"; nvidia-smi --query-gpu=index,utilization.gpu,memory.used --format=csv
