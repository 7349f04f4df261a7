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
chk inplace_mc FSB_XCHG_MODE=inplace FSB_XCHG_MULTICAST=1
chk inplace_peer FSB_XCHG_MODE=inplace FSB_XCHG_MULTICAST=0
if [ "$N" = "2" ]; then chk gather_peer FSB_XCHG_MODE=gather FSB_XCHG_MULTICAST=0; fi
run peer FSB_EXCHANGE=peer
if [ "$N" != "8" ]; then run nccl FSB_EXCHANGE=nccl; fi
echo "elapsed ${SECONDS}s"
