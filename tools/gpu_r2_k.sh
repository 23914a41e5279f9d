#!/bin/bash
O=/root/repo/gpurun_out/r2k
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
for shape in "f32 36928 3072 3 1" "f32 36928 768 10 0" "f32 2560 768 20 0" "f32 2560 3072 10 1" "f32 1000 768 0 0" "f32 333 203 0 1"; do
  timeout 120 $S f64 $shape >> $O/f64_cases.log 2>&1
done
cat $O/f64_cases.log
for shape in "768 3072 5" "3072 768 5" "2304 768 5" "768 768 5"; do timeout 120 $S rhs $shape >> $O/rhs.log 2>&1; done
cat $O/rhs.log
timeout 900 python -m pytest tests/test_gpu_merge.py tests/test_gpu_regmean_chain.py tests/test_gpu_large.py -q -x 2>&1 | tail -3
python tools/time_regmean.py base 2>&1 | tail -4
