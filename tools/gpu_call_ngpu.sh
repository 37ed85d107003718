#!/bin/bash
# usage: gpu_call_ngpu.sh N   -- N-rank bit-identity check + bench at N GPUs (weak C2 + strong C3)
N=$1
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/r2_topo_${N}gpu.txt 2>&1
nproc > $OUT/r2_cpus_${N}gpu.txt; free -g >> $OUT/r2_cpus_${N}gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_check.py 2>&1 | grep -E "MGPU|rc=" > $OUT/r2_mgpu$N.log; echo "mgpu rc=${PIPESTATUS[0]}" >> $OUT/r2_mgpu$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2_bench_n$N.json 2> $OUT/r2_bench_n$N.err; echo "bench rc=$?" >> $OUT/r2_bench_n$N.err
tail -4 $OUT/r2_mgpu$N.log
grep -E "rc=" $OUT/r2_bench_n$N.err | tail -2
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads([l for l in open('gpurun_out/r2_bench_n%s.json'%N) if l.startswith('{')][-1]); st=d['stages']
print('N',N,'ms %.3f value %.0f'%(d['ms_per_step'], d['value']), st['launch'][:40], 'ab %.3f fuse %.3f exch %.3f'%(st['stages_ab_ms_per_step'], st.get('exchange_plus_fuse_ms_per_step',0), st['exchange_ms_per_step']), 'mgpu', d.get('mgpu_bit_identical'), 'e2e', d['e2e']['value'])
c=d['c3']; print('c3 ms', c['ms_per_step'], 'value', c['value'], 'ab', c['stages']['stages_ab_ms_per_step'], 'xf', c['stages']['exchange_plus_fuse_ms_per_step'], 'frac', c['frac_of_hbm_peak'], c['nvlink'])
PY
