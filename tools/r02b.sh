#!/bin/bash
# round-2 GPU visit b: ncu --set full of the cfg4 compositing kernels (baseline before the rewrite) and of the
# sort / projection / Adam kernels (VERDICT item 6 asks for their summaries)
TAG=r02b; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
export FSB_PRUNE_LISTS=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'raster_(bwd|seg|fold|stop)_kernel' --launch-skip 84 -c 10 \
   -o $OUT/${TAG}_raster_cfg4 -f python tools/stage_bench.py cfg4 2 > $OUT/${TAG}_ncu_raster.log 2>&1
echo "raster t=${SECONDS}s"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'onesweep|radix_hist|project_sh_bwd|project_sh_fwd|adam_multi|isect_reach|isect_emit|ssim' --launch-skip 120 -c 24 \
   -o $OUT/${TAG}_other_cfg4 -f python tools/stage_bench.py cfg4 2 > $OUT/${TAG}_ncu_other.log 2>&1
echo "other t=${SECONDS}s"
ls -la $OUT/${TAG}*
