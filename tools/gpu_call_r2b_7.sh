#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sparse.py -m gpu -x -q -k "fusion or sparse" > $OUT/r2b_pytest7.log 2>&1; tail -2 $OUT/r2b_pytest7.log
timeout 600 python tools/microbench.py fuse 2>&1 | tee $OUT/r2b_microbench7a.txt | grep -i "fuse V= *[1-4]00"
VISSAT_MB_SHUFFLE=1 timeout 600 python tools/microbench.py fuse 2>&1 | tee $OUT/r2b_microbench7b.txt | grep -i "fuse V= *[1-4]00"
