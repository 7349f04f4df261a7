#!/bin/bash
# compute-sanitizer over the kernels with cross-thread / cross-kernel protocols: memcheck and racecheck on
# the compositing kernels (mbarrier + bulk-copy staging, per-unit workspace shared between forward kernels), the
# onesweep sort (look-back status words) and the captured step's overflow path.  Small cases only: the tools slow the
# kernels 10-100x.   Usage (through gpurun): bash tools/sanitize.sh [tag]
TAG=${1:-r02}; OUT=gpurun_out; mkdir -p $OUT
SEL='tests/test_gpu_raster_dn.py::test_flagged_entries_reach_set_b_only tests/test_gpu_render.py::test_radix_sort_pairs_stable[8-4097] tests/test_gpu_render.py::test_radix_sort_pairs_stable[45-100003]'
RAS='tests/test_gpu_render.py -k raster_forward_backward'
for TOOL in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 7 --target-processes all \
     python -m pytest $SEL -m gpu -q -x > $OUT/${TAG}_sanitizer_${TOOL}.log 2>&1
  echo "$TOOL exit $?" >> $OUT/${TAG}_sanitizer_${TOOL}.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit" $OUT/${TAG}_sanitizer_${TOOL}.log | tail -6
done
