#!/bin/bash
# round-2 GPU visit d: full GPU suite on the rewritten compositing path, bench.py (cfg4 default + cfg2 secondary),
# ncu --set full of the new compositing kernels on cfg4
TAG=r02d; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -8 $OUT/${TAG}_pytest_gpu.log; echo "t=${SECONDS}s"
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; head -c 600 $OUT/${TAG}_bench.json; echo; tail -5 $OUT/${TAG}_bench.err; echo "t=${SECONDS}s"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'raster_|unit_table|tile_flag' --launch-skip 100 -c 8 \
   -o $OUT/${TAG}_raster_cfg4 -f python tools/stage_bench.py cfg4 2 > $OUT/${TAG}_ncu_raster.log 2>&1
echo "ncu t=${SECONDS}s"
timeout 300 python tools/stage_bench.py cfg2 20 > $OUT/${TAG}_stage_cfg2.json 2> $OUT/${TAG}_stage_cfg2.err; cat $OUT/${TAG}_stage_cfg2.json; tail -3 $OUT/${TAG}_stage_cfg2.err
echo "elapsed ${SECONDS}s"
