#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sparse.py tests/test_gpu_fullsize.py tests/test_gpu_run_fuse.py tests/test_aggregate_3d.py -m gpu -q -s > $OUT/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_pytest5.log
timeout 900 python bench.py --no-cpu-baseline > $OUT/r2_bench5.json 2> $OUT/r2_bench5.err; echo "bench rc=$?" >> $OUT/r2_bench5.err
VISSAT_K2_LEGACY=1 timeout 600 python bench.py --no-cpu-baseline --no-c3 > $OUT/r2_bench5_legacyk2.json 2>> $OUT/r2_bench5.err
timeout 300 python tools/microbench.py > $OUT/r2_microbench5.txt 2>&1
grep -E "passed|failed|rc=" $OUT/r2_pytest5.log | tail -3; tail -2 $OUT/r2_bench5.err
