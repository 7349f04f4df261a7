#!/bin/bash
# Short GPU visit: the tests named on the command line, then per-stage kernel times of cfg2.
TAG=${1:-r01i}; shift
OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 300 python -m pytest "$@" -m gpu -q --tb=short > $OUT/${TAG}_pytest_quick.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_quick.log
grep -E "^(FAILED|ERROR)|passed|failed|pytest exit" $OUT/${TAG}_pytest_quick.log | tail -12; echo "t=${SECONDS}s"
timeout 200 python tools/stage_bench.py cfg2 20 > $OUT/${TAG}_stage_cfg2.json 2> $OUT/${TAG}_stage_cfg2.err; cat $OUT/${TAG}_stage_cfg2.json; tail -3 $OUT/${TAG}_stage_cfg2.err
echo "elapsed ${SECONDS}s"
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_graph.json 2> $OUT/${TAG}_bench_graph.err; python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_graph.json")); print("bench", d["value"], d["e2e"]["value"], d["gpu_launches_per_step"], d["roofline"]["kernel_ms_all"])
except Exception as e:
    print("bench failed", e); print(open("$OUT/${TAG}_bench_graph.err").read()[-1500:])
PY
echo "elapsed ${SECONDS}s"
