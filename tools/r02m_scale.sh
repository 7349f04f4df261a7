#!/bin/bash
# round-2 scaling visit: bench.py at N GPUs as the driver launches it; N=8 also with NVSwitch multicast and with NCCL
N=${1:-8}; TAG=r02m; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 420 $TR bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline --no-secondary > $OUT/${TAG}_bench_n${N}_${name}.json 2> $OUT/${TAG}_bench_n${N}_${name}.err
  echo "rc=$? $name t=${SECONDS}s"; head -c 330 $OUT/${TAG}_bench_n${N}_${name}.json; echo; grep -v "^\s*$" $OUT/${TAG}_bench_n${N}_${name}.err | tail -2 | cut -c1-300
}
run peer FSB_EXCHANGE=peer
run nccl FSB_EXCHANGE=nccl
if [ "$N" = "8" ]; then
  run peer_mc FSB_EXCHANGE=peer FSB_XCHG_MULTICAST=1
  timeout 300 $TR tools/peer_exchange_check.py 1000000 > $OUT/${TAG}_xchg_check_n${N}.json 2> $OUT/${TAG}_xchg_check_n${N}.err; echo "rc=$?"; tail -1 $OUT/${TAG}_xchg_check_n${N}.json | cut -c1-600
  FSB_XCHG_MULTICAST=1 timeout 300 $TR tools/peer_exchange_check.py 1000000 > $OUT/${TAG}_xchg_check_n${N}_mc.json 2> $OUT/${TAG}_xchg_check_n${N}_mc.err; echo "rc=$?"; tail -1 $OUT/${TAG}_xchg_check_n${N}_mc.json | cut -c1-600
fi
echo "elapsed ${SECONDS}s"
