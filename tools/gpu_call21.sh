#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_parity.py -m gpu -q > $OUT/r2_pytest21.log 2>&1; tail -3 $OUT/r2_pytest21.log
timeout 900 python bench.py --no-cpu-baseline > $OUT/r2_bench21.json 2> $OUT/r2_bench21.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench21.json') if l.startswith('{')][-1]); st=d['stages']
print('C2 ms %.3f ab %.3f k1 %.1f k2 %.1f fuse %.3f'%(d['ms_per_step'], st['stages_ab_ms_per_step'], st['k1_isolated_ms_per_view']*1e3, st['k2_isolated_ms_per_view']*1e3, st['k3_fuse_ms_per_step']))
c=d['c3']; s3=c['stages']; print('C3 ms %.2f ab %.2f k1 %.1f k2 %.1f fuse %.2f frac %.3f'%(c['ms_per_step'], s3['stages_ab_ms_per_step'], s3['k1_isolated_ms_per_view']*1e3, s3['k2_isolated_ms_per_view']*1e3, s3['k3_fuse_ms_per_step'], c['frac_of_hbm_peak']))
PY
for c in C5; do timeout 600 python bench.py --config $c --no-cpu-baseline --no-c3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); st=d['stages']
print('$c %.1f Gpix/s %.3f ms k1 %.1f k2 %.1f'%(d['value']/1e3, d['ms_per_step'], st['k1_isolated_ms_per_view']*1e3, st['k2_isolated_ms_per_view']*1e3))"; done
