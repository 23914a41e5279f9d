#!/bin/bash
O=/root/repo/gpurun_out/r2ag
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
timeout 900 python -m pytest tests/test_gpu_gram.py -q 2>&1 | tail -4
for a in "36928 768 10 0" "36928 3072 10 1"; do timeout 120 $S i8x4 $a 2>&1 | grep -E "I8X4|FAIL|error"; done
