#!/bin/bash
# round-2 GPU check C: fp64 DMMA SYRK, chain test in all modes, selftest
O=/root/repo/gpurun_out/r2c
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
timeout 900 $S quick > $O/selftest_quick.log 2>&1; echo "selftest exit $?" >> $O/selftest_quick.log
grep -E "F64|SPLIT|FAIL|exit|launches" $O/selftest_quick.log
timeout 900 python -m pytest tests/test_gpu_regmean_chain.py -q -s > $O/chain.log 2>&1; echo "chain exit $?" >> $O/chain.log
grep -E "regmean errors|passed|failed|Error|error|exit" $O/chain.log | head -40
for shape in "f32 36928 3072 3 1" "f32 36928 768 10 0" "bf16 36928 3072 3 1" "f32 2560 768 20 0" "f32 2560 3072 10 1"; do
  timeout 120 $S f64 $shape >> $O/f64_cases.log 2>&1
done
cat $O/f64_cases.log
timeout 120 $S case f32 36928 768 20 0 >> $O/cases.log 2>&1
timeout 120 $S case f32 36928 3072 10 1 >> $O/cases.log 2>&1
cat $O/cases.log
