#!/bin/bash
# torchrun --no-python tools/ncu_rank0.sh <metrics> <tag> : rank 0 under ncu (stage-B launches only), the other ranks plain
METRICS=$1; TAG=$2
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --metrics $METRICS --clock-control none -k regex:k_grid_finalize -s 4 -c 8 --csv --log-file gpurun_out/nvl_${TAG}_rank0.csv python tools/nvlink_probe.py
else
  exec python tools/nvlink_probe.py
fi
