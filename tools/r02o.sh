#!/bin/bash
# round-2 GPU visit o: full tests after the pack / unit-table / loss changes, ncu of projection backward, Adam, losses
# and the two-pixel backward (traffic JSON for bench.py), torch glue profile
TAG=r02o; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -12 $OUT/${TAG}_pytest_gpu.log | cut -c1-300; cp $OUT/parity_metrics.json $OUT/${TAG}_parity_metrics.json; echo "t=${SECONDS}s"
timeout 300 python tools/stage_bench.py cfg4 10 > $OUT/${TAG}_stage_cfg4.json 2> $OUT/${TAG}_stage_cfg4.err; cat $OUT/${TAG}_stage_cfg4.json; tail -2 $OUT/${TAG}_stage_cfg4.err
timeout 300 python tools/stage_bench.py cfg2 20 > $OUT/${TAG}_stage_cfg2.json 2> $OUT/${TAG}_stage_cfg2.err; cat $OUT/${TAG}_stage_cfg2.json
echo "t=${SECONDS}s"
timeout 900 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; head -c 400 $OUT/${TAG}_bench_default.json; echo; tail -3 $OUT/${TAG}_bench_default.err | cut -c1-300
echo "t=${SECONDS}s"
timeout 300 python tools/glue_profile.py cfg4 > $OUT/${TAG}_glue_cfg4.txt 2> $OUT/${TAG}_glue_cfg4.err; head -30 $OUT/${TAG}_glue_cfg4.txt | cut -c1-200; tail -2 $OUT/${TAG}_glue_cfg4.err | cut -c1-300
echo "t=${SECONDS}s"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'project_sh_bwd|adam_multi|dn_loss|raster_bwd2|ssim_fwd|ssim_bwd|densify_stats|normals_bwd|radix_hist' --launch-skip 24 -c 16 \
   -o $OUT/${TAG}_bwdside_cfg4 -f python tools/stage_bench.py cfg4 2 > $OUT/${TAG}_ncu.log 2>&1
tail -3 $OUT/${TAG}_ncu.log | cut -c1-300
echo "elapsed ${SECONDS}s"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'vh_' -c 6 -o $OUT/${TAG}_hull -f python tools/hull_bench.py 512 1 > $OUT/${TAG}_ncu_hull.log 2>&1
tail -3 $OUT/${TAG}_ncu_hull.log | cut -c1-300
echo "elapsed ${SECONDS}s"
