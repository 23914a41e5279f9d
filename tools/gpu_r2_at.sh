#!/bin/bash
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out/r2at
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_gramfile.py -q -x 2>&1 | tail -2
for mode in "" "--symmetric"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tools/exchange_trace.py $mode 2>&1 | grep -v "OMP\|\*\*\*" | tee -a gpurun_out/r2at/trace.log | tail -8
done
