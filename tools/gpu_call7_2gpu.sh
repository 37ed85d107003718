#!/bin/bash
# 2 GPUs: N-rank bit-identity (both transports, sparse peer exchange) + bench at N = 2 (weak C2 + strong C3)
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/r2_topo_2gpu.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_check.py > $OUT/r2_mgpu2.log 2>&1; echo "mgpu rc=$?" >> $OUT/r2_mgpu2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/r2_bench_n2.json 2> $OUT/r2_bench_n2.err; echo "bench rc=$?" >> $OUT/r2_bench_n2.err
tail -12 $OUT/r2_mgpu2.log; tail -5 $OUT/r2_bench_n2.err | cut -c1-400
