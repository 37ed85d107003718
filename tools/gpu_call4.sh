#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -q -s > $OUT/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_pytest4.log
# C3 on one GPU: per-launch times of the fusion kernels (sparse path), one step
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum --clock-control none -k regex:'k_fuse|k_median' --csv \
    --log-file $OUT/launches_c3_fuse_r2a.csv python bench.py --config C3 --steps 1 --warmup 0 --no-cpu-baseline --no-c3 > $OUT/ncu_c3_r2a.log 2>&1
tail -3 $OUT/r2_pytest4.log; tail -2 $OUT/ncu_c3_r2a.log | cut -c1-300
