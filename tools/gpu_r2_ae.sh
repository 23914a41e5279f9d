#!/bin/bash
# int8x4: two epilogue slabs
O=/root/repo/gpurun_out/r2ae
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
for a in "64 128 0 0" "1000 768 0 0" "333 384 0 1" "2560 768 10 0" "2560 3072 10 1" "36928 768 10 0" "36928 3072 10 1" "36928 4096 5 0" "36928 3072 5 1 f16"; do
  timeout 120 $S i8x4 $a 2>&1 | grep -E "I8X4|FAIL|error" | tee -a $O/cases.log
done
timeout 900 python -m pytest tests/test_gpu_gram.py tests/test_gpu_regmean_chain.py -q 2>&1 | tail -2
for tool in memcheck synccheck racecheck; do timeout 300 compute-sanitizer --tool $tool $S i8x4 1000 768 0 0 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY"; done
