#!/bin/bash
# round-2 GPU visit l: shared-memory reduce-scatter in both backward kernels, per-Gaussian prepack records,
# launch list of graph replays (cfg4), ncu of the non-compositing kernels
TAG=r02l; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -12 $OUT/${TAG}_pytest_gpu.log | cut -c1-300; cp $OUT/parity_metrics.json $OUT/${TAG}_parity_metrics.json; echo "t=${SECONDS}s"
timeout 300 python tools/stage_bench.py cfg4 10 > $OUT/${TAG}_stage_cfg4.json 2> $OUT/${TAG}_stage_cfg4.err; cat $OUT/${TAG}_stage_cfg4.json; tail -2 $OUT/${TAG}_stage_cfg4.err
timeout 300 python tools/stage_bench.py cfg2 20 > $OUT/${TAG}_stage_cfg2.json 2> $OUT/${TAG}_stage_cfg2.err; cat $OUT/${TAG}_stage_cfg2.json
echo "t=${SECONDS}s"
timeout 900 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; head -c 400 $OUT/${TAG}_bench_default.json; echo; tail -3 $OUT/${TAG}_bench_default.err | cut -c1-300
echo "t=${SECONDS}s"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
   --log-file $OUT/${TAG}_launches_graph_cfg4.csv env FSB_PROFILE=cfg4 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/${TAG}_ncu_launch.log 2>&1
tail -2 $OUT/${TAG}_ncu_launch.log | cut -c1-200
echo "t=${SECONDS}s"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'project_sh|adam_multi|dn_loss|compose|normal|flatness|loss_combine|raster_prepack|raster_pack|raster_fwd' --launch-skip 60 -c 24 \
   -o $OUT/${TAG}_other_cfg4 -f python tools/stage_bench.py cfg4 2 > $OUT/${TAG}_ncu.log 2>&1
tail -3 $OUT/${TAG}_ncu.log | cut -c1-300
echo "elapsed ${SECONDS}s"
