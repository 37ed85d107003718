#!/bin/bash
# K1 magic floor A/B is vs the previous call's numbers (76.1 Gpix/s, k1iso 31.5); fusion pair vs large at V=200
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or coscheduled or exact_mode" > $OUT/r2b_pytest9.log 2>&1; tail -2 $OUT/r2b_pytest9.log
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --no-c3 --no-e2e ${CFG:+--config $CFG} 2>$OUT/r2b_err_$label.log | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); st=d['stages']
print('$label ${CFG:-C2} %.1f Gpix/s %.3f ms  ab %.3f ms (%.1f us/view) k1iso %.1f k2iso %.1f fuse %.3f  launches %d'%(d['value']/1e3, d['ms_per_step'], st['stages_ab_ms_per_step'], st['stages_ab_effective_ms_per_view']*1e3, st['k1_isolated_ms_per_view']*1e3, st['k2_isolated_ms_per_view']*1e3, st.get('k3_fuse_ms_per_step',0), d['gpu_launches']))" || tail -3 $OUT/r2b_err_$label.log
}
run magic X=1
run magic2 X=1
CFG=C4 run magic X=1
CFG=C5 run magic X=1
VISSAT_MB_SHUFFLE=1 VISSAT_FUSE_PAIR=0 timeout 600 python tools/microbench.py fuse 2>&1 | grep -i "fuse V= *[1-4]00" | sed 's/^/PAIR=0 shuffled /'
VISSAT_FUSE_PAIR=0 timeout 600 python tools/microbench.py fuse 2>&1 | grep -i "fuse V= *[1-4]00" | sed 's/^/PAIR=0 /'
