#!/bin/bash
# round-2 GPU visit q: KNN after the box-growth rewrite (tests, 1M-point measurement), e2e A/B of the 8-bit targets
TAG=r02q; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 600 python -m pytest tests/test_knn.py tests/test_seed_points.py tests/test_gpu_graph_step.py -m gpu -q --maxfail=40 > $OUT/${TAG}_pytest_knn.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_knn.log
tail -8 $OUT/${TAG}_pytest_knn.log | cut -c1-300; echo "t=${SECONDS}s"
timeout 200 python tools/knn_bench.py 1000000 16 surface > $OUT/${TAG}_knn_surface.json 2> $OUT/${TAG}_knn_surface.err; cat $OUT/${TAG}_knn_surface.json; tail -2 $OUT/${TAG}_knn_surface.err | cut -c1-300
timeout 200 python tools/knn_bench.py 1000000 16 uniform > $OUT/${TAG}_knn_uniform.json 2> $OUT/${TAG}_knn_uniform.err; cat $OUT/${TAG}_knn_uniform.json; tail -2 $OUT/${TAG}_knn_uniform.err | cut -c1-300
echo "t=${SECONDS}s"
timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 300 > $OUT/${TAG}_bench_u8.json 2> $OUT/${TAG}_bench_u8.err; head -c 1500 $OUT/${TAG}_bench_u8.json | grep -o '"value": [0-9.]*\|"h2d_bytes_per_step": [0-9]*' | head -4; tail -2 $OUT/${TAG}_bench_u8.err | cut -c1-200
FSB_U8_TARGETS=0 timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 300 > $OUT/${TAG}_bench_f32.json 2> $OUT/${TAG}_bench_f32.err; head -c 1500 $OUT/${TAG}_bench_f32.json | grep -o '"value": [0-9.]*\|"h2d_bytes_per_step": [0-9]*' | head -4; tail -2 $OUT/${TAG}_bench_f32.err | cut -c1-200
timeout 300 python bench.py --config cfg2 --no-cpu-baseline --no-secondary --steps 1000 > $OUT/${TAG}_bench_cfg2_u8.json 2> $OUT/${TAG}_bench_cfg2_u8.err; head -c 1500 $OUT/${TAG}_bench_cfg2_u8.json | grep -o '"value": [0-9.]*\|"h2d_bytes_per_step": [0-9]*' | head -4
FSB_U8_TARGETS=0 timeout 300 python bench.py --config cfg2 --no-cpu-baseline --no-secondary --steps 1000 > $OUT/${TAG}_bench_cfg2_f32.json 2> $OUT/${TAG}_bench_cfg2_f32.err; head -c 1500 $OUT/${TAG}_bench_cfg2_f32.json | grep -o '"value": [0-9.]*\|"h2d_bytes_per_step": [0-9]*' | head -4
echo "t=${SECONDS}s"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'knn_query|knn_cells|knn_cell_start|knn_gather|gaussian_density' -c 5 \
   -o $OUT/${TAG}_knn -f python tools/knn_bench.py 1000000 16 uniform > $OUT/${TAG}_ncu_knn.log 2>&1
tail -2 $OUT/${TAG}_ncu_knn.log | cut -c1-300
echo "elapsed ${SECONDS}s"
