#!/bin/bash
# ncu captures of the hot path (run under gpurun, 1 GPU).  Outputs land in gpurun_out/.
#   tools/profile.sh <tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
# every launch of our kernels with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv \
    --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/ncu_launches_$TAG.log 2>&1
# full sets: K1 + K2 of the timed step (skip the 102 launches of the warm-up step), then fusion + blur
ncu --set full --clock-control none --import-source on -k regex:'k_unproject_scatter|k_grid_finalize' -s 102 -c 4 \
    -f -o $OUT/prof_k1k2_$TAG $BENCH > $OUT/ncu_k1k2_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_fuse|k_median3x3' -s 2 -c 2 \
    -f -o $OUT/prof_fuse_$TAG $BENCH > $OUT/ncu_fuse_$TAG.log 2>&1
ls -la $OUT
