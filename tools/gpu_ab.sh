#!/bin/bash
# GPU visit for an A/B of DNSplatterStep.fused_outputs: full parity suite (all failures listed), bench with the switch
# on and off, launch list of the fused step.
TAG=${1:-r01h}
OUT=gpurun_out
mkdir -p $OUT
SECONDS=0
timeout 600 python -m pytest tests -m gpu -q --tb=short --durations=5 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|pytest exit" $OUT/${TAG}_pytest_gpu.log | tail -15; echo "t=${SECONDS}s"
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 400 python bench.py --steps 50 --warmup 5 > $OUT/${TAG}_bench_graph.json 2> $OUT/${TAG}_bench_graph.err; python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_graph.json")); print("ON ", d["value"], d["e2e"]["value"], d["gpu_launches_per_step"], d["roofline"]["kernel_ms_all"])
except Exception as e:
    print("bench ON failed", e); print(open("$OUT/${TAG}_bench_graph.err").read()[-1500:])
PY
echo "t=${SECONDS}s"
FSB_FUSED_OUTPUTS=0 timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_graph_unfused.json 2> $OUT/${TAG}_bench_graph_unfused.err; python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_graph_unfused.json")); print("OFF", d["value"], d["e2e"]["value"], d["gpu_launches_per_step"])
except Exception as e:
    print("bench OFF failed", e)
PY
echo "t=${SECONDS}s"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
   --log-file $OUT/${TAG}_launches_graph.csv env FSB_PROFILE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
echo "elapsed ${SECONDS}s"
