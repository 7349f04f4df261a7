#!/bin/bash
# round-2 GPU visit j: SSIM register blocking, backward dot-product form, hull byte votes + reciprocal, bench step window
TAG=r02j; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -8 $OUT/${TAG}_pytest_gpu.log; cp $OUT/parity_metrics.json $OUT/${TAG}_parity_metrics.json; echo "t=${SECONDS}s"
timeout 300 python tools/stage_bench.py cfg4 10 > $OUT/${TAG}_stage_cfg4.json 2> $OUT/${TAG}_stage_cfg4.err; cat $OUT/${TAG}_stage_cfg4.json; tail -2 $OUT/${TAG}_stage_cfg4.err
timeout 300 python tools/stage_bench.py cfg2 20 > $OUT/${TAG}_stage_cfg2.json 2> $OUT/${TAG}_stage_cfg2.err; cat $OUT/${TAG}_stage_cfg2.json
timeout 300 python tools/hull_bench.py 512 5 > $OUT/${TAG}_hull_n1.json 2> $OUT/${TAG}_hull_n1.err; cat $OUT/${TAG}_hull_n1.json | cut -c1-700; tail -2 $OUT/${TAG}_hull_n1.err
echo "t=${SECONDS}s"
timeout 900 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; head -c 400 $OUT/${TAG}_bench_default.json; echo; tail -3 $OUT/${TAG}_bench_default.err | cut -c1-300
echo "t=${SECONDS}s"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'onesweep|radix_hist|raster_bwd2|ssim|isect_reach|vh_' --launch-skip 60 -c 16 \
   -o $OUT/${TAG}_misc_cfg4 -f python tools/stage_bench.py cfg4 2 > $OUT/${TAG}_ncu_misc.log 2>&1
echo "elapsed ${SECONDS}s"
