#!/bin/bash
# key-grid clear folded into stage B (two alternating key grids per stream): parity + A/B timings
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_run_fuse.py tests/test_gpu_sparse.py -m gpu -x -q > $OUT/r2b_pytest4.log 2>&1; tail -3 $OUT/r2b_pytest4.log
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --no-c3 --no-e2e ${CFG:+--config $CFG} 2>$OUT/r2b_err_$label.log | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); st=d['stages']
print('$label ${CFG:-C2} %.1f Gpix/s %.3f ms  ab %.3f ms (%.1f us/view) k1iso %.1f k2iso %.1f fuse %.3f  launches %d'%(d['value']/1e3, d['ms_per_step'], st['stages_ab_ms_per_step'], st['stages_ab_effective_ms_per_view']*1e3, st['k1_isolated_ms_per_view']*1e3, st['k2_isolated_ms_per_view']*1e3, st.get('k3_fuse_ms_per_step',0), d['gpu_launches']))" || tail -3 $OUT/r2b_err_$label.log
}
run memset VISSAT_FOLD_CLEAR=0
run fold VISSAT_FOLD_CLEAR=1
run fold_s3 VISSAT_FOLD_CLEAR=1 VISSAT_STREAMS=3
run fold_k1c2 VISSAT_FOLD_CLEAR=1 VISSAT_K1_CTAS_PER_SM=2
for CFG in C1 C4 C5; do run memset VISSAT_FOLD_CLEAR=0; run fold VISSAT_FOLD_CLEAR=1; done
