#!/bin/bash
# N GPUs: the bench under torchrun exactly as the driver launches it
N=${1:-8}
O=/root/repo/gpurun_out/r2g
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 ) > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench exit $?" >> $O/bench_n$N.err
tail -6 $O/bench_n$N.err
python tools/show_bench.py $O/bench_n$N.json | cut -c1-1200
