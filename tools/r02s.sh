#!/bin/bash
# round-2 GPU visit s (final check): full tests, smoke, default bench (shared graph pool, host-time split of the
# refinement leg) and its A/B, KNN measurement with the fixed timer
TAG=r02s; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -8 $OUT/${TAG}_pytest_gpu.log | cut -c1-300; cp $OUT/parity_metrics.json $OUT/${TAG}_parity_metrics.json; echo "t=${SECONDS}s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log | cut -c1-200
timeout 900 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; head -c 400 $OUT/${TAG}_bench_default.json; echo; grep -o '"with_refinement": {[^}]*}[^}]*}' $OUT/${TAG}_bench_default.json | cut -c1-700; tail -3 $OUT/${TAG}_bench_default.err | cut -c1-300
echo "t=${SECONDS}s"
timeout 900 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_default_pool0.json 2> $OUT/${TAG}_bench_default_pool0.err; head -c 300 $OUT/${TAG}_bench_default_pool0.json; echo; grep -o '"with_refinement": {[^}]*}[^}]*}' $OUT/${TAG}_bench_default_pool0.json | cut -c1-700
echo "t=${SECONDS}s"
timeout 200 python tools/knn_bench.py 1000000 16 surface > $OUT/${TAG}_knn_surface.json 2> $OUT/${TAG}_knn_surface.err; cat $OUT/${TAG}_knn_surface.json; tail -2 $OUT/${TAG}_knn_surface.err | cut -c1-300
timeout 200 python tools/knn_bench.py 1000000 16 uniform > $OUT/${TAG}_knn_uniform.json 2> $OUT/${TAG}_knn_uniform.err; cat $OUT/${TAG}_knn_uniform.json; tail -2 $OUT/${TAG}_knn_uniform.err | cut -c1-300
echo "elapsed ${SECONDS}s"
