#!/bin/bash
# 2 GPUs: NCCL parity test + the bench under torchrun
O=/root/repo/gpurun_out/r2f
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_large.py -q -rs > $O/dist_test.log 2>&1; echo "exit $?" >> $O/dist_test.log
tail -8 $O/dist_test.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench exit $?" >> $O/bench_n2.err
tail -5 $O/bench_n2.err
python tools/show_bench.py $O/bench_n2.json | cut -c1-1500
