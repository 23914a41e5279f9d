#!/bin/bash
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out/r2ap
for mode in "" "--symmetric"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tools/exchange_trace.py $mode 2>&1 | grep -v "OMP\|\*\*\*" | tee -a gpurun_out/r2ap/trace.log | tail -9
done
