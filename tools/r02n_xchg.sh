#!/bin/bash
# round-2: in-place all-reduce (NVLS multimem.ld_reduce + multimem.st / peer loads + stores) at N GPUs
N=${1:-2}; TAG=r02n; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
chk() { name=$1; shift
  env "$@" timeout 240 $TR tools/peer_exchange_check.py 1000000 > $OUT/${TAG}_xchg_n${N}_${name}.json 2> $OUT/${TAG}_xchg_n${N}_${name}.err; echo "rc=$? $name t=${SECONDS}s"; tail -1 $OUT/${TAG}_xchg_n${N}_${name}.json | cut -c1-700; grep -v "^\s*$" $OUT/${TAG}_xchg_n${N}_${name}.err | tail -2 | cut -c1-300; }
run() { name=$1; shift
  env "$@" timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > $OUT/${TAG}_bench_n${N}_${name}.json 2> $OUT/${TAG}_bench_n${N}_${name}.err
  echo "rc=$? $name t=${SECONDS}s"; head -c 330 $OUT/${TAG}_bench_n${N}_${name}.json; echo; grep -v "^\s*$" $OUT/${TAG}_bench_n${N}_${name}.err | tail -2 | cut -c1-300; }
if [ "$N" = "2" ]; then
  chk inplace_peer_k4 FSB_XCHG_MODE=inplace FSB_XCHG_MULTICAST=0 FSB_XCHG_CHUNKS=4
  chk inplace_mc_k4 FSB_XCHG_MODE=inplace FSB_XCHG_MULTICAST=1 FSB_XCHG_CHUNKS=4
  chk inplace_peer_k2 FSB_XCHG_MODE=inplace FSB_XCHG_MULTICAST=0 FSB_XCHG_CHUNKS=2
  chk inplace_peer_k1 FSB_XCHG_MODE=inplace FSB_XCHG_MULTICAST=0 FSB_XCHG_CHUNKS=1
  run peer_inplace_k4 FSB_EXCHANGE=peer FSB_XCHG_MODE=inplace
else
  chk inplace_mc_k4 FSB_XCHG_CHUNKS=4
  chk inplace_mc_k8 FSB_XCHG_CHUNKS=8
  run peer_k4 FSB_EXCHANGE=peer FSB_XCHG_CHUNKS=4
  run peer_k8 FSB_EXCHANGE=peer FSB_XCHG_CHUNKS=8
  if [ "$N" != "8" ]; then run nccl FSB_EXCHANGE=nccl; fi
fi
echo "elapsed ${SECONDS}s"
