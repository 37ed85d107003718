#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c3 --no-e2e"
ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_unproject_scatter|k_grid_finalize' -s 102 -c 2 \
    -f -o $OUT/prof_k1k2_r2c $BENCH > $OUT/ncu_k1k2_r2c.log 2>&1
tail -2 $OUT/ncu_k1k2_r2c.log
