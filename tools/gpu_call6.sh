#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:k_grid_finalize_keys -s 5 -c 1 -o $OUT/prof_k2keys_r2 python tools/microbench.py k2 > $OUT/ncu_k2keys.log 2>&1
VISSAT_K2_LEGACY=1 timeout 300 $NCU -k regex:k_grid_finalize -s 5 -c 1 -o $OUT/prof_k2legacy_r2 python tools/microbench.py k2 > $OUT/ncu_k2legacy.log 2>&1
timeout 400 $NCU -k regex:k_fuse_pair -s 1 -c 1 -o $OUT/prof_pair_r2 python tools/microbench.py fuse > $OUT/ncu_pair.log 2>&1
VISSAT_FUSE_PAIR=0 timeout 400 $NCU -k regex:'k_fuse_large' -s 1 -c 1 -o $OUT/prof_large_r2 python tools/microbench.py fuse > $OUT/ncu_large.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_fuse|k_median' -c 20 --csv \
    --log-file $OUT/launches_c3_fuse_r2b.csv python bench.py --config C3 --steps 1 --warmup 0 --no-cpu-baseline --no-c3 > $OUT/ncu_c3_r2b.log 2>&1
ls -la $OUT/*.ncu-rep
