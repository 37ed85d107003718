#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
N=$1
nvidia-smi topo -m > $OUT/r2b_topo_${N}gpu.txt 2>&1
( python tools/pcie_probe.py; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 tools/pcie_probe.py ) 2>&1 | grep -E "H2D|D2H|both" | tee $OUT/r2b_pcie_probe_n$N.txt
