#!/bin/bash
# round-2 GPU visit x (last check): full tests incl. the level-set kernel, smoke
TAG=r02x; OUT=gpurun_out; mkdir -p $OUT
SECONDS=0
timeout 100 python -m pytest tests/test_level_set.py tests/test_knn.py -m gpu -q -x > $OUT/${TAG}_pytest_f2.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_f2.log; tail -25 $OUT/${TAG}_pytest_f2.log | cut -c1-250; echo "t=${SECONDS}s"
timeout 600 python -m pytest tests -m gpu -q --maxfail=40 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -6 $OUT/${TAG}_pytest_gpu.log | cut -c1-300; echo "t=${SECONDS}s"
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log | cut -c1-200
echo "elapsed ${SECONDS}s"
