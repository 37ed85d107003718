#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
( time python bench.py --impl reference --steps 20 --warmup 5 ) > $OUT/r2_bench_reference.json 2> $OUT/r2_bench_reference.err
for c in C1 C4 C5; do timeout 600 python bench.py --config $c --no-cpu-baseline --no-c3 > $OUT/r2_bench_n1_$c.json 2>/dev/null; done
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_reference.json') if l.startswith('{')][-1])
print('reference arm: %.3f Mpix/s, pass %.1f s (views %.1f s, fusion %.1f s), measured_not_extrapolated %s, cores %s'%(d['value'], d['ms_per_step']/1e3, d['t_views_s'], d['t_fuse_s'], d['measured_not_extrapolated'], d['cpu_baseline']['cores']))
for c in ('C1','C4','C5'):
    d=json.loads([l for l in open('gpurun_out/r2_bench_n1_%s.json'%c) if l.startswith('{')][-1]); st=d['stages']
    print(c, '%.1f Gpix/s %.3f ms'%(d['value']/1e3, d['ms_per_step']), 'k1 %.1f k2 %.1f fuse %.3f'%(st['k1_isolated_ms_per_view']*1e3, st['k2_isolated_ms_per_view']*1e3, st.get('k3_fuse_ms_per_step',0)), st['launch'][:30])
PY
tail -4 $OUT/r2_bench_reference.err
