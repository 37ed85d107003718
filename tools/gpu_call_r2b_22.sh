#!/bin/bash
# final check of round 2: full GPU test-suite, smoke, default bench, reference arm
OUT=gpurun_out
mkdir -p $OUT
( time timeout 2400 python -m pytest tests -m gpu -q -s ) > $OUT/r2b_pytest_final.log 2>&1; echo "pytest rc=$?" >> $OUT/r2b_pytest_final.log
grep -E "passed|failed|FAILED|rc=|real" $OUT/r2b_pytest_final.log | tail -6
( time timeout 300 python __graft_entry__.py smoke ) > $OUT/r2b_smoke_final.log 2>&1; tail -2 $OUT/r2b_smoke_final.log | head -1
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $OUT/r2b_bench_final.json 2> $OUT/r2b_bench_final.err; tail -c 400 $OUT/r2b_bench_final.json; tail -3 $OUT/r2b_bench_final.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $OUT/r2b_bench_reference_final.json 2> $OUT/r2b_bench_reference_final.err; tail -c 600 $OUT/r2b_bench_reference_final.json; tail -3 $OUT/r2b_bench_reference_final.err
