#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_geodesy_pin.py tests/test_gpu_parity.py -m gpu -q -s > $OUT/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_pytest3.log
timeout 900 python bench.py > $OUT/r2_bench3.json 2> $OUT/r2_bench3.err; echo "bench rc=$?" >> $OUT/r2_bench3.err
tail -5 $OUT/r2_pytest3.log; tail -3 $OUT/r2_bench3.err
