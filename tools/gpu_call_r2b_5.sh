#!/bin/bash
# split-search selection in the multi-lane fusion kernels: parity + timings
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sparse.py -m gpu -x -q -k "fusion or sparse" > $OUT/r2b_pytest5.log 2>&1; tail -3 $OUT/r2b_pytest5.log
timeout 600 python tools/microbench.py fuse 2>&1 | tee $OUT/r2b_microbench5.txt | grep -i fuse
