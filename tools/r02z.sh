#!/bin/bash
# round-2 GPU visit z: f2 tests after the empty-neighbour-slot guards
OUT=gpurun_out; mkdir -p $OUT
timeout 100 python -m pytest tests/test_level_set.py tests/test_knn.py tests/test_abi_symbols.py -q -x > $OUT/r02z_pytest_f2.log 2>&1; echo "exit $?" >> $OUT/r02z_pytest_f2.log; tail -4 $OUT/r02z_pytest_f2.log | cut -c1-250
