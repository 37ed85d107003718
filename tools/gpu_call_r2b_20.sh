#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/nvl_*.csv "$OUT/nvl_%p.csv"
run() {
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 --no-python \
    tools/ncu_rank0.sh $1 $2 > $OUT/nvlink_probe_$2.log 2>&1
  echo "$2 rc=$?"; grep -E "NVLINK_PROBE|ERROR" $OUT/nvlink_probe_$2.log | head -5
  [ -f $OUT/nvl_$2_rank0.csv ] && grep -E "nvl|gpu__time" $OUT/nvl_$2_rank0.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | head -40
}
run gpu__time_duration.sum time 29651
run nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum nvlink 29652
