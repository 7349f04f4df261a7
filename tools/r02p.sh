#!/bin/bash
# round-2 GPU visit p: full tests (KNN, seed points, u8 targets, flat projection backward), projection backward A/B,
# default bench, KNN / seed-cloud measurements, ncu of the new kernels
TAG=r02p; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -15 $OUT/${TAG}_pytest_gpu.log | cut -c1-300; cp $OUT/parity_metrics.json $OUT/${TAG}_parity_metrics.json; echo "t=${SECONDS}s"
timeout 300 python tools/stage_bench.py cfg4 10 > $OUT/${TAG}_stage_cfg4.json 2> $OUT/${TAG}_stage_cfg4.err; cat $OUT/${TAG}_stage_cfg4.json; tail -2 $OUT/${TAG}_stage_cfg4.err
FSB_PROJ_BWD_FLAT=0 timeout 300 python tools/stage_bench.py cfg4 10 > $OUT/${TAG}_stage_cfg4_flat0.json 2> $OUT/${TAG}_stage_cfg4_flat0.err; cat $OUT/${TAG}_stage_cfg4_flat0.json
timeout 300 python tools/stage_bench.py cfg2 20 > $OUT/${TAG}_stage_cfg2.json 2> $OUT/${TAG}_stage_cfg2.err; cat $OUT/${TAG}_stage_cfg2.json
echo "t=${SECONDS}s"
timeout 900 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; head -c 400 $OUT/${TAG}_bench_default.json; echo; tail -3 $OUT/${TAG}_bench_default.err | cut -c1-300
echo "t=${SECONDS}s"
timeout 300 python tools/knn_bench.py 1000000 16 surface > $OUT/${TAG}_knn_surface.json 2> $OUT/${TAG}_knn_surface.err; cat $OUT/${TAG}_knn_surface.json; tail -2 $OUT/${TAG}_knn_surface.err | cut -c1-300
timeout 300 python tools/knn_bench.py 1000000 16 uniform > $OUT/${TAG}_knn_uniform.json 2> $OUT/${TAG}_knn_uniform.err; cat $OUT/${TAG}_knn_uniform.json; tail -2 $OUT/${TAG}_knn_uniform.err | cut -c1-300
timeout 300 python tools/seed_bench.py > $OUT/${TAG}_seed.json 2> $OUT/${TAG}_seed.err; cat $OUT/${TAG}_seed.json; tail -2 $OUT/${TAG}_seed.err | cut -c1-300
echo "t=${SECONDS}s"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'project_sh_bwd' --launch-skip 2 -c 2 \
   -o $OUT/${TAG}_projbwd_cfg4 -f python tools/stage_bench.py cfg4 2 > $OUT/${TAG}_ncu.log 2>&1
tail -2 $OUT/${TAG}_ncu.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'knn_|gaussian_density' -c 12 \
   -o $OUT/${TAG}_knn -f python tools/knn_bench.py 1000000 16 surface > $OUT/${TAG}_ncu_knn.log 2>&1
tail -2 $OUT/${TAG}_ncu_knn.log | cut -c1-300
echo "elapsed ${SECONDS}s"
