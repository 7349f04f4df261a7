#!/bin/bash
# round-2 first GPU visit: validate the pruned-list path and A/B it on cfg2 (graph) and cfg4 (eager stages)
TAG=r02a; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
FSB_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_prune_lists.py -m gpu -q > $OUT/${TAG}_pytest_prune.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_prune.log
tail -15 $OUT/${TAG}_pytest_prune.log; echo "t=${SECONDS}s"
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_cfg2.json 2> $OUT/${TAG}_bench_cfg2.err; head -c 400 $OUT/${TAG}_bench_cfg2.json; echo
FSB_PRUNE_LISTS=1 timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_cfg2_prune.json 2> $OUT/${TAG}_bench_cfg2_prune.err; head -c 400 $OUT/${TAG}_bench_cfg2_prune.json; echo; tail -3 $OUT/${TAG}_bench_cfg2_prune.err
echo "t=${SECONDS}s"
timeout 300 python tools/stage_bench.py cfg4 10 > $OUT/${TAG}_stage_cfg4.json 2> $OUT/${TAG}_stage_cfg4.err; cat $OUT/${TAG}_stage_cfg4.json
FSB_PRUNE_LISTS=1 timeout 300 python tools/stage_bench.py cfg4 10 > $OUT/${TAG}_stage_cfg4_prune.json 2> $OUT/${TAG}_stage_cfg4_prune.err; cat $OUT/${TAG}_stage_cfg4_prune.json; tail -3 $OUT/${TAG}_stage_cfg4_prune.err
echo "elapsed ${SECONDS}s"
