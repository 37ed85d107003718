#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
N=$1
VISSAT_PROBE_VIEWS=50 VISSAT_PROBE_STEPS=20 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29661 tools/nvlink_probe.py 2>&1 | grep -E "NVLINK_|Error|error" | tee $OUT/r2b_nvlink_counters_n$N.txt
