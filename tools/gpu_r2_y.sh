#!/bin/bash
# 2 GPUs: NCCL parity tests (incl. the packed fp64 exchange of the RegMean-grade cache) and the N = 2 bench line
O=/root/repo/gpurun_out/r2y
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_large.py -q 2>&1 | tee $O/pytest_2gpu.log | tail -5
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; tail -3 $O/bench_n2.err
python tools/show_bench.py $O/bench_n2.json 2>/dev/null | head -8
