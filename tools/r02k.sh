#!/bin/bash
# round-2 GPU visit k: two-level binning (depth sort of the Gaussians, then a stable sort on the tile bits only)
TAG=r02k; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_prune_lists.py tests/test_gpu_raster_dn.py tests/test_gpu_graph_step.py tests/test_gpu_render.py tests/test_gpu_step.py -m gpu -q --maxfail=40 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -12 $OUT/${TAG}_pytest_gpu.log | cut -c1-300; echo "t=${SECONDS}s"
for tl in 1 0; do
FSB_TWO_LEVEL_BINNING=$tl timeout 300 python tools/stage_bench.py cfg4 10 > $OUT/${TAG}_stage_cfg4_tl$tl.json 2> $OUT/${TAG}_stage_cfg4_tl$tl.err; cat $OUT/${TAG}_stage_cfg4_tl$tl.json; tail -2 $OUT/${TAG}_stage_cfg4_tl$tl.err
FSB_TWO_LEVEL_BINNING=$tl timeout 300 python tools/stage_bench.py cfg2 20 > $OUT/${TAG}_stage_cfg2_tl$tl.json 2> $OUT/${TAG}_stage_cfg2_tl$tl.err; cat $OUT/${TAG}_stage_cfg2_tl$tl.json
done
FSB_RASTER_BWD_REDUCE=shfl timeout 300 python tools/stage_bench.py cfg4 10 > $OUT/${TAG}_stage_cfg4_shfl.json 2> $OUT/${TAG}_stage_cfg4_shfl.err; cat $OUT/${TAG}_stage_cfg4_shfl.json; tail -2 $OUT/${TAG}_stage_cfg4_shfl.err
echo "t=${SECONDS}s"
timeout 900 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; head -c 400 $OUT/${TAG}_bench_default.json; echo; tail -3 $OUT/${TAG}_bench_default.err | cut -c1-300
echo "t=${SECONDS}s"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'onesweep|radix_hist|isect_reach|scan_|isect_offsets|raster_pack|unit_table|tile_flag|raster_bwd2' --launch-skip 90 -c 27 \
   -o $OUT/${TAG}_binning_cfg4 -f python tools/stage_bench.py cfg4 2 > $OUT/${TAG}_ncu.log 2>&1
tail -3 $OUT/${TAG}_ncu.log | cut -c1-300
echo "elapsed ${SECONDS}s"
