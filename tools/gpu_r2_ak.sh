#!/bin/bash
O=/root/repo/gpurun_out/r2ak
mkdir -p $O
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/symm_probe.py 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tee $O/probe.log | tail -30
nvidia-smi topo -m 2>&1 | head -12
