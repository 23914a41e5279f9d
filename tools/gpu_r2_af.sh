#!/bin/bash
O=/root/repo/gpurun_out/r2af
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
timeout 300 compute-sanitizer --tool racecheck --racecheck-report all $S i8x4 1000 768 0 0 > $O/race_i8.log 2>&1
timeout 300 compute-sanitizer --tool racecheck --racecheck-report all $S case f32 1000 768 0 0 > $O/race_tf32.log 2>&1
grep -E "hazard|Race|RACECHECK|at .*\+0x|syrk_|\.cu" $O/race_i8.log | head -40
echo ----
grep -E "RACECHECK SUMMARY" $O/race_tf32.log
