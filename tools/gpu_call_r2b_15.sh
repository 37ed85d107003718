#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
VISSAT_MB_BASE=50 ncu --set full --clock-control none --import-source on -k regex:'k_fuse_large' -s 2 -c 1 -f -o $OUT/prof_large2_r2b python tools/microbench.py fuse > $OUT/ncu_large2_r2b.log 2>&1
tail -3 $OUT/ncu_large2_r2b.log
timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-c3 | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print('C2 %.1f Gpix/s e2e %s'%(d['value']/1e3, d['e2e']))"
