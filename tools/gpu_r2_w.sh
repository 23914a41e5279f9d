#!/bin/bash
# int8x4: phase-split schedule (A/B), 16-bit activations, full selftest, gram tests, default bench line
O=/root/repo/gpurun_out/r2w
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
for ps in 1 0; do export VLM_I8_PHASE_SPLIT=$ps; echo "phase split: $ps" | tee -a $O/cases.log
for a in "2560 768 10 0" "36928 768 10 0" "36928 1024 10 0" "9216 768 10 0" "36928 3072 5 1"; do
  timeout 120 $S i8x4 $a 2>&1 | grep -E "I8X4|FAIL|error" | tee -a $O/cases.log
done
done
unset VLM_I8_PHASE_SPLIT
for a in "36928 3072 5 1 f16" "36928 768 10 0 bf16" "333 384 0 1 f16"; do
  timeout 120 $S i8x4 $a 2>&1 | grep -E "I8X4|FAIL|error" | tee -a $O/cases.log
done
timeout 600 $S quick > $O/selftest_quick.log 2>&1; tail -3 $O/selftest_quick.log; grep -c OK $O/selftest_quick.log; grep FAIL $O/selftest_quick.log
timeout 900 python -m pytest tests/test_gpu_gram.py tests/test_gpu_regmean_chain.py tests/test_cabi.py -q 2>&1 | tail -3
timeout 1500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -2 $O/bench_n1.err
python tools/show_bench.py $O/bench_n1.json 2>/dev/null | head -60
