#!/bin/bash
# round-2 GPU visit h: two-pixels-per-lane backward (A/B against the one-pixel kernel), faster reach mask, 5-pass sort
TAG=r02h; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -8 $OUT/${TAG}_pytest_gpu.log; cp $OUT/parity_metrics.json $OUT/${TAG}_parity_metrics.json; echo "t=${SECONDS}s"
for PX in 0 1; do
  FSB_RASTER_BWD_PX=$PX timeout 300 python tools/stage_bench.py cfg4 10 > $OUT/${TAG}_stage_cfg4_px$PX.json 2> $OUT/${TAG}_stage_cfg4_px$PX.err; cat $OUT/${TAG}_stage_cfg4_px$PX.json; tail -2 $OUT/${TAG}_stage_cfg4_px$PX.err
  FSB_RASTER_BWD_PX=$PX timeout 300 python tools/stage_bench.py cfg2 20 > $OUT/${TAG}_stage_cfg2_px$PX.json 2> $OUT/${TAG}_stage_cfg2_px$PX.err; cat $OUT/${TAG}_stage_cfg2_px$PX.json
done
echo "t=${SECONDS}s"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'raster_|unit_table|tile_flag' --launch-skip 100 -c 8 \
   -o $OUT/${TAG}_raster_cfg4 -f python tools/stage_bench.py cfg4 2 > $OUT/${TAG}_ncu_raster.log 2>&1
echo "ncu t=${SECONDS}s"
bash tools/sanitize.sh $TAG
echo "elapsed ${SECONDS}s"
