#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
VISSAT_MB_SHUFFLE=1 timeout 600 python tools/microbench.py fuse 2>&1 | tee $OUT/r2b_microbench6.txt | grep -i "fuse V= *[1-4]00"
