#!/bin/bash
# C3 on one GPU: internal stream count
OUT=gpurun_out
mkdir -p $OUT
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 600 python bench.py --config C3 --steps 3 --warmup 1 --no-cpu-baseline --no-c3 --no-e2e 2>$OUT/r2b_err_$label.log | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); st=d['stages']
print('$label C3 %.1f Gpix/s %.2f ms  ab %.2f ms (%.1f us/view) k1 %.1f k2 %.1f k1iso %.1f k2iso %.1f fuse %.2f  launches %d'%(d['value']/1e3, d['ms_per_step'], st['stages_ab_ms_per_step'], st['stages_ab_effective_ms_per_view']*1e3, st['k1_unproject_scatter_ms_per_view']*1e3, st['k2_grid_finalize_ms_per_view']*1e3, st['k1_isolated_ms_per_view']*1e3, st['k2_isolated_ms_per_view']*1e3, st.get('k3_fuse_ms_per_step',0), d['gpu_launches']))" || tail -3 $OUT/r2b_err_$label.log
}
run s4 VISSAT_STREAMS=4
run s3 VISSAT_STREAMS=3
run s2 VISSAT_STREAMS=2
run s1 VISSAT_STREAMS=1
timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-c3 --no-e2e | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); st=d['stages']
print('C2 %.1f Gpix/s %.3f ms k2iso %.1f'%(d['value']/1e3, d['ms_per_step'], st['k2_isolated_ms_per_view']*1e3))"
