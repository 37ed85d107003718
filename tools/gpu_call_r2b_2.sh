#!/bin/bash
# co-scheduled stage A+B kernel: parity test, then A/B timings on C2 (and C4 / C5 / C1 shapes)
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coscheduled or captured_step or end_to_end" > $OUT/r2b_pytest2.log 2>&1; tail -3 $OUT/r2b_pytest2.log
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --no-c3 --no-e2e ${CFG:+--config $CFG} 2>$OUT/r2b_err_$label.log | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); st=d['stages']
print('$label ${CFG:-C2} %.1f Gpix/s %.3f ms  ab %.3f ms (%.1f us/view) fuse %.3f  launches %d'%(d['value']/1e3, d['ms_per_step'], st['stages_ab_ms_per_step'], st['stages_ab_effective_ms_per_view']*1e3, st.get('k3_fuse_ms_per_step',0), d['gpu_launches']))" || tail -3 $OUT/r2b_err_$label.log
}
run off VISSAT_AB=0
run ab_s2 VISSAT_AB=1
run ab_s1 VISSAT_AB_STREAMS=1
run ab_s3 VISSAT_AB_STREAMS=3
run ab_s4 VISSAT_AB_STREAMS=4
run ab_c2 VISSAT_AB_CTAS_PER_SM=2
for CFG in C1 C4 C5; do run off VISSAT_AB=0; run ab VISSAT_AB=1; done
