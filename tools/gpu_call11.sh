#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --no-cpu-baseline --no-c3 --steps 10 --warmup 3"
for k in 2 4; do VISSAT_K1_CTAS_PER_SM=$k timeout 300 $B > $OUT/r2_exp_k1ctas$k.json 2>/dev/null; done
for s in 2 3; do VISSAT_STREAMS=$s timeout 300 $B > $OUT/r2_exp_streams$s.json 2>/dev/null; done
VISSAT_K1_CTAS_PER_SM=2 VISSAT_STREAMS=3 timeout 300 $B > $OUT/r2_exp_k1ctas2_streams3.json 2>/dev/null
timeout 900 python bench.py --no-cpu-baseline > $OUT/r2_bench11.json 2> $OUT/r2_bench11.err
timeout 300 python tools/microbench.py fuse > $OUT/r2_microbench11.txt 2>&1
for f in $OUT/r2_exp_*.json $OUT/r2_bench11.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); st=d['stages']
    print(sys.argv[1].split('/')[-1], 'ms %.3f'%d['ms_per_step'], 'ab %.3f'%st['stages_ab_ms_per_step'], 'k1iso %.1f k2iso %.1f'%(st['k1_isolated_ms_per_view']*1e3, st['k2_isolated_ms_per_view']*1e3), 'c3' , (d.get('c3') or {}).get('ms_per_step'), ((d.get('c3') or {}).get('stages') or {}).get('k3_fuse_ms_per_step'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
cat $OUT/r2_microbench11.txt | tail -12
