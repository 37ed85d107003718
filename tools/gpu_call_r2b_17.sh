#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
for p in 1 2; do VISSAT_PRIO=$p timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or captured_step or coscheduled" 2>&1 | tail -1 | sed "s/^/PRIO=$p /"; done
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --no-c3 --no-e2e ${CFG:+--config $CFG} 2>$OUT/r2b_err_$label.log | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); st=d['stages']
print('$label ${CFG:-C2} %.1f Gpix/s %.3f ms  ab %.3f ms (%.1f us/view) k1 %.1f k2 %.1f fuse %.3f'%(d['value']/1e3, d['ms_per_step'], st['stages_ab_ms_per_step'], st['stages_ab_effective_ms_per_view']*1e3, st['k1_unproject_scatter_ms_per_view']*1e3, st['k2_grid_finalize_ms_per_view']*1e3, st.get('k3_fuse_ms_per_step',0)))" || tail -3 $OUT/r2b_err_$label.log
}
run prio0 VISSAT_PRIO=0
run prio1 VISSAT_PRIO=1
run prio2 VISSAT_PRIO=2
run prio1_s3 VISSAT_PRIO=1 VISSAT_STREAMS=3
run prio2_s3 VISSAT_PRIO=2 VISSAT_STREAMS=3
run prio1_s2 VISSAT_PRIO=1 VISSAT_STREAMS=2
for CFG in C4 C5; do run prio0 VISSAT_PRIO=0; run prio1 VISSAT_PRIO=1; run prio2 VISSAT_PRIO=2; done
