#!/bin/bash
# full validation: smoke, GPU test suite, native selftest
O=/root/repo/gpurun_out/r2z
mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tee $O/gpu_tests.log | tail -5
S=vl-merging_b200/csrc/build/selftest
for a in "2560 768 10 0" "9248 256 10 0" "36928 768 10 0"; do timeout 120 $S i8x4 $a 2>&1 | grep -E "I8X4|FAIL|error"; done
