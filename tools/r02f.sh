#!/bin/bash
# round-2 GPU visit f (2 GPUs): own NVLink gradient exchange: correctness tool, then bench N=2 with both exchanges
TAG=r02f; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
timeout 300 $TR tools/peer_exchange_check.py 1000000 > $OUT/${TAG}_xchg_check.json 2> $OUT/${TAG}_xchg_check.err; echo "rc=$?"; cat $OUT/${TAG}_xchg_check.json; tail -5 $OUT/${TAG}_xchg_check.err | cut -c1-400; echo "t=${SECONDS}s"
FSB_XCHG_MULTICAST=1 timeout 300 $TR tools/peer_exchange_check.py 1000000 > $OUT/${TAG}_xchg_check_mc.json 2> $OUT/${TAG}_xchg_check_mc.err; echo "rc=$?"; cat $OUT/${TAG}_xchg_check_mc.json; tail -5 $OUT/${TAG}_xchg_check_mc.err | cut -c1-400; echo "t=${SECONDS}s"
FSB_EXCHANGE=peer timeout 400 $TR bench.py --gpus 2 --steps 50 --warmup 5 > $OUT/${TAG}_bench_n2_peer.json 2> $OUT/${TAG}_bench_n2_peer.err; echo "rc=$?"; head -c 300 $OUT/${TAG}_bench_n2_peer.json; echo; tail -3 $OUT/${TAG}_bench_n2_peer.err | cut -c1-300; echo "t=${SECONDS}s"
timeout 400 $TR bench.py --gpus 2 --steps 50 --warmup 5 > $OUT/${TAG}_bench_n2_nccl.json 2> $OUT/${TAG}_bench_n2_nccl.err; echo "rc=$?"; head -c 300 $OUT/${TAG}_bench_n2_nccl.json; echo; tail -3 $OUT/${TAG}_bench_n2_nccl.err | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_graph_step.py tests/test_reference_dn_model.py -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; tail -5 $OUT/${TAG}_pytest.log
echo "elapsed ${SECONDS}s"; nvidia-smi --query-gpu=index,utilization.gpu,memory.used --format=csv
