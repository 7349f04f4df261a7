#!/bin/bash
# round-2 GPU visit i: analytic reach mask, identity unit order, bwd kernel heuristic, metrics shim, eval render loop;
# full default bench.py (with the CPU sample) for its wall time
TAG=r02i; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -8 $OUT/${TAG}_pytest_gpu.log; cp $OUT/parity_metrics.json $OUT/${TAG}_parity_metrics.json; echo "t=${SECONDS}s"
timeout 300 python tools/stage_bench.py cfg4 10 > $OUT/${TAG}_stage_cfg4.json 2> $OUT/${TAG}_stage_cfg4.err; cat $OUT/${TAG}_stage_cfg4.json; tail -2 $OUT/${TAG}_stage_cfg4.err
timeout 300 python tools/stage_bench.py cfg2 20 > $OUT/${TAG}_stage_cfg2.json 2> $OUT/${TAG}_stage_cfg2.err; cat $OUT/${TAG}_stage_cfg2.json
echo "t=${SECONDS}s"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'raster_pack|unit_table|raster_fwd_kernel' --launch-skip 60 -c 6 \
   -o $OUT/${TAG}_pack_cfg4 -f python tools/stage_bench.py cfg4 2 > $OUT/${TAG}_ncu_pack.log 2>&1
echo "ncu t=${SECONDS}s"
( time timeout 1200 python bench.py ) > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; head -c 400 $OUT/${TAG}_bench_default.json; echo; tail -4 $OUT/${TAG}_bench_default.err
echo "t=${SECONDS}s"
FSB_STEP_METRICS=1 timeout 600 python bench.py --config cfg2 --steps 500 --no-cpu-baseline > $OUT/${TAG}_bench_cfg2_metrics.json 2> $OUT/${TAG}_bench_cfg2_metrics.err; head -c 300 $OUT/${TAG}_bench_cfg2_metrics.json; echo
timeout 600 python bench.py --config cfg2 --steps 500 --no-cpu-baseline > $OUT/${TAG}_bench_cfg2.json 2> $OUT/${TAG}_bench_cfg2.err; head -c 300 $OUT/${TAG}_bench_cfg2.json; echo
echo "elapsed ${SECONDS}s"
