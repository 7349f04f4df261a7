#!/bin/bash
# round-2 GPU visit c: first run of the rewritten compositing kernels (packed records + bulk copies, units, one walk
# for both colour sets): parity first, then timings
TAG=r02c; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_raster_dn.py tests/test_gpu_render.py -m gpu -q --maxfail=40 -k "not radix and not projection" > $OUT/${TAG}_pytest_raster.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_raster.log
tail -40 $OUT/${TAG}_pytest_raster.log; echo "t=${SECONDS}s"
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --deselect tests/test_gpu_raster_dn.py --deselect tests/test_gpu_render.py > $OUT/${TAG}_pytest_rest.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_rest.log
tail -30 $OUT/${TAG}_pytest_rest.log; echo "t=${SECONDS}s"
timeout 300 python tools/stage_bench.py cfg4 10 > $OUT/${TAG}_stage_cfg4.json 2> $OUT/${TAG}_stage_cfg4.err; cat $OUT/${TAG}_stage_cfg4.json; tail -3 $OUT/${TAG}_stage_cfg4.err
timeout 300 python tools/stage_bench.py cfg2 20 > $OUT/${TAG}_stage_cfg2.json 2> $OUT/${TAG}_stage_cfg2.err; cat $OUT/${TAG}_stage_cfg2.json; tail -3 $OUT/${TAG}_stage_cfg2.err
echo "elapsed ${SECONDS}s"
