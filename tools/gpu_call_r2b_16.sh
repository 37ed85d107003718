#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sparse.py -m gpu -x -q -k "fusion or sparse" > $OUT/r2b_pytest16.log 2>&1; tail -2 $OUT/r2b_pytest16.log
VISSAT_MB_BASE=50 timeout 600 python tools/microbench.py fuse 2>&1 | grep -i "fuse V= *[1-4]00" | tee $OUT/r2b_microbench16.txt
timeout 900 python bench.py --config C3 --steps 3 --warmup 1 --no-cpu-baseline --no-c3 --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); st=d['stages']
print('C3 %.1f Gpix/s %.2f ms  ab %.2f ms fuse %.2f'%(d['value']/1e3, d['ms_per_step'], st['stages_ab_ms_per_step'], st.get('k3_fuse_ms_per_step',0)))"
