#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
for b in 0 1; do for mode in "VISSAT_MB_SHUFFLE=0 VISSAT_MB_BASE=4" "VISSAT_MB_SHUFFLE=1 VISSAT_MB_BASE=4" "VISSAT_MB_SHUFFLE=0 VISSAT_MB_BASE=50"; do
  env VISSAT_FUSE_BISECT=$b $mode timeout 600 python tools/microbench.py fuse 2>&1 | grep -i "fuse V= *400" | sed "s/^/bisect=$b $mode: /"
done; done | tee $OUT/r2b_microbench13.txt
