#!/bin/bash
# ncu: the co-scheduled kernel (one mid-step launch with both roles) -- stall reasons and pipe utilisation
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c3 --no-e2e"
VISSAT_AB=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_ab -s 12 -c 2 \
    -f -o $OUT/prof_ab_r2b $BENCH > $OUT/ncu_ab_r2b.log 2>&1
tail -3 $OUT/ncu_ab_r2b.log
ncu -i $OUT/prof_ab_r2b.ncu-rep --page raw --csv > $OUT/prof_ab_r2b_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/prof_ab_r2b_raw.csv')))
hdr=rows[0]; 
want=['gpu__time_duration.sum','smsp__inst_executed.sum','sm__inst_executed_pipe_alu','sm__inst_executed_pipe_fma','sm__pipe_alu_cycles_active','smsp__issue_active.avg.pct','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__average_warp','smsp__warp_issue_stalled','launch__registers','sm__throughput','sm__inst_executed_pipe_xu','pipe_fp64','l1tex__data_bank_conflicts']
for i,h in enumerate(hdr):
    if any(w in h for w in want):
        print(h, [r[i] for r in rows[2:4]])
PY
