#!/bin/bash
# final state of round 2 (session 2): full GPU test-suite, bench records, ncu launch list + full captures
OUT=gpurun_out
mkdir -p $OUT
( time timeout 2400 python -m pytest tests -m gpu -q -s ) > $OUT/r2b_pytest10.log 2>&1; echo "pytest rc=$?" >> $OUT/r2b_pytest10.log
grep -E "^\[|passed|failed|FAILED|rc=|real" $OUT/r2b_pytest10.log | tail -30
( time timeout 300 python __graft_entry__.py smoke ) > $OUT/r2b_smoke10.log 2>&1; tail -2 $OUT/r2b_smoke10.log
( time timeout 1200 python bench.py ) > $OUT/r2b_bench_n1_c2_c3.json 2> $OUT/r2b_bench_n1.err; tail -c 300 $OUT/r2b_bench_n1_c2_c3.json; tail -2 $OUT/r2b_bench_n1.err
for c in C1 C4 C5; do timeout 600 python bench.py --config $c --no-cpu-baseline --no-c3 > $OUT/r2b_bench_n1_$c.json 2>/dev/null; done
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c3 --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv \
    --log-file $OUT/launches_r2b.csv $BENCH > $OUT/ncu_launches_r2b.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_unproject_scatter|k_grid_finalize' -s 102 -c 4 \
    -f -o $OUT/prof_k1k2_r2b $BENCH > $OUT/ncu_k1k2_r2b.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_fuse|k_median3x3' -s 2 -c 2 \
    -f -o $OUT/prof_fuse_r2b $BENCH > $OUT/ncu_fuse_r2b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_grid_finalize' -s 51 -c 1 \
    -f -o $OUT/prof_k2cold_r2b $BENCH > $OUT/ncu_k2cold_r2b.log 2>&1
VISSAT_MB_SHUFFLE=1 ncu --set full --clock-control none -k regex:'k_fuse_large' -s 2 -c 1 -f -o $OUT/prof_large_r2b python tools/microbench.py fuse > $OUT/ncu_large_r2b.log 2>&1
ls -la $OUT | tail -30
