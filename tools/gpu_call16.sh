#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_parity.py tests/test_gpu_fullconfig.py -m gpu -q -s > $OUT/r2_pytest16.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_pytest16.log
grep -E "^\[|passed|failed|FAILED|rc=" $OUT/r2_pytest16.log | tail -20
timeout 900 python bench.py --no-cpu-baseline > $OUT/r2_bench16.json 2> $OUT/r2_bench16.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench16.json') if l.startswith('{')][-1]); st=d['stages']
print('C2 ms %.3f ab %.3f k1 %.1f k2 %.1f fuse %.3f'%(d['ms_per_step'], st['stages_ab_ms_per_step'], st['k1_isolated_ms_per_view']*1e3, st['k2_isolated_ms_per_view']*1e3, st['k3_fuse_ms_per_step']))
c=d['c3']; s3=c['stages']; print('C3 ms %.2f ab %.2f k1 %.1f k2 %.1f fuse %.2f frac %.3f'%(c['ms_per_step'], s3['stages_ab_ms_per_step'], s3['k1_isolated_ms_per_view']*1e3, s3['k2_isolated_ms_per_view']*1e3, s3['k3_fuse_ms_per_step'], c['frac_of_hbm_peak']))
PY
echo "--- K1 warp-aggregated scatter A/B"; timeout 300 python tools/microbench.py k1 2>&1 | grep "^k1"; VISSAT_K1_WARPAGG=1 timeout 300 python tools/microbench.py k1 2>&1 | grep "^k1" | sed 's/^/WARPAGG /'
