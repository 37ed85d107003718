#!/bin/bash
# GPU call 1 (round 2): new parity tests + baseline bench + warm-cache ncu captures.
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv > $OUT/r2_gpuinfo.txt; nproc >> $OUT/r2_gpuinfo.txt; free -g >> $OUT/r2_gpuinfo.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_pytest1.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/r2_bench1.json 2> $OUT/r2_bench1.err
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
# warm-cache (no flush between kernels, as in the pipelined step) captures of K1/K2 and fusion
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_unproject_scatter|k_grid_finalize' -s 102 -c 4 \
    -f -o $OUT/prof_k1k2_r2warm $BENCH > $OUT/ncu_k1k2_r2warm.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_fuse|k_median3x3' -s 2 -c 2 \
    -f -o $OUT/prof_fuse_r2warm $BENCH > $OUT/ncu_fuse_r2warm.log 2>&1
tail -5 $OUT/r2_pytest1.log
