#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (our arm + reference arm), ncu launch list, one --set full capture of
# the raster kernels (raw + source pages), cfg4 stage times.
# Usage (through gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_smi.txt 2>&1
SECONDS=0; timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -3 $OUT/${TAG}_pytest_gpu.log; echo "t=${SECONDS}s"
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > $OUT/${TAG}_bench_graph.json 2> $OUT/${TAG}_bench_graph.err; tail -c 300 $OUT/${TAG}_bench_graph.json; echo "t=${SECONDS}s"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; tail -c 300 $OUT/${TAG}_bench_reference.json; echo "t=${SECONDS}s"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
   --log-file $OUT/${TAG}_launches_graph.csv env FSB_PROFILE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'raster_(bwd|seg)_kernel' -c 6 \
   -o $OUT/${TAG}_raster_full -f env FSB_PROFILE=1 python bench.py --steps 1 --warmup 3 --mode eager --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ncu -i $OUT/${TAG}_raster_full.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_raster.csv 2>/dev/null
ncu -i $OUT/${TAG}_raster_full.ncu-rep --page source --csv -k regex:'raster_bwd_kernel' > $OUT/${TAG}_ncu_source_raster_bwd.csv 2>/dev/null
echo "t=${SECONDS}s"
timeout 300 python tools/stage_bench.py cfg4 10 > $OUT/${TAG}_stage_cfg4.json 2> $OUT/${TAG}_stage_cfg4.err; tail -c 400 $OUT/${TAG}_stage_cfg4.json
echo "elapsed ${SECONDS}s"; ls -la $OUT | tail -14
