#!/bin/bash
# grouped launches with 64-chunk minimum pieces
O=/root/repo/gpurun_out/r2ai
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
$S batch 9 36928 768 0 0 0 10 2>&1 | tail -1 | tee -a $O/batch.log
$S batch 36 2560 768 12 2560 3072 10 2>&1 | tail -1 | tee -a $O/batch.log
$S batch 24 2560 768 8 2560 3072 10 2>&1 | tail -1 | tee -a $O/batch.log
$S batch 3 2560 768 1 2560 3072 10 2>&1 | tail -1 | tee -a $O/batch.log
timeout 600 python -m pytest tests/test_gpu_gram.py tests/test_gpu_fused.py -q 2>&1 | tail -2
timeout 900 python bench.py --no-variants --no-vitl --no-irtr --no-cpu-baseline --no-gpu-baseline --no-gramfile --no-regmean --steps 20 > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
python tools/show_bench.py $O/bench.json | head -5
