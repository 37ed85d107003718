#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/r2_bench_n2b.json 2> $OUT/r2_bench_n2b.err; echo "bench rc=$?" >> $OUT/r2_bench_n2b.err
grep -E "Error|error|rc=" $OUT/r2_bench_n2b.err | tail -5
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n2b.json') if l.startswith('{')][-1]); st=d['stages']
print('N2 ms %.3f value %.0f'%(d['ms_per_step'], d['value']), st['launch'], 'ab %.3f'%st['stages_ab_ms_per_step'], 'mgpu', d.get('mgpu_bit_identical'))
c=d['c3']; print('c3 ms', c['ms_per_step'], c['stages']['launch'], c['stages']['stages_ab_ms_per_step'], c['stages']['exchange_plus_fuse_ms_per_step'], c['nvlink'])
PY
