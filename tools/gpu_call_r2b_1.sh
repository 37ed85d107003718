#!/bin/bash
# round 2, session 2, call 1: state check -- full GPU test-suite, smoke, default bench
OUT=gpurun_out
mkdir -p $OUT
( time timeout 2400 python -m pytest tests -m gpu -x -q -s ) > $OUT/r2b_pytest1.log 2>&1; echo "pytest rc=$?" >> $OUT/r2b_pytest1.log
grep -E "^\[|passed|failed|FAILED|rc=|real" $OUT/r2b_pytest1.log | tail -30
( time timeout 300 python __graft_entry__.py smoke ) > $OUT/r2b_smoke1.log 2>&1; tail -3 $OUT/r2b_smoke1.log
( time timeout 1200 python bench.py ) > $OUT/r2b_bench1.json 2> $OUT/r2b_bench1.err; tail -c 600 $OUT/r2b_bench1.json; tail -3 $OUT/r2b_bench1.err
