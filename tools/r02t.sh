#!/bin/bash
# round-2 GPU visit t: default bench after the e2e warm-up fix (graph pool change reverted), graph-step tests
TAG=r02t; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 300 python -m pytest tests/test_gpu_graph_step.py tests/test_gpu_metrics.py -m gpu -q > $OUT/${TAG}_pytest_graph.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_graph.log; tail -3 $OUT/${TAG}_pytest_graph.log | cut -c1-200
timeout 900 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; head -c 400 $OUT/${TAG}_bench_default.json; echo; grep -o '"e2e": {[^}]*}' $OUT/${TAG}_bench_default.json; grep -o '"with_refinement": {[^}]*}[^}]*}' $OUT/${TAG}_bench_default.json | cut -c1-500; tail -3 $OUT/${TAG}_bench_default.err | cut -c1-300
echo "elapsed ${SECONDS}s"
