#!/bin/bash
O=/root/repo/gpurun_out/r2o
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
timeout 300 $S simtopk 130 300 192 0 > $O/simtopk.log 2>&1
timeout 300 $S simtopk 5000 25000 768 10 >> $O/simtopk.log 2>&1
timeout 300 $S simtopk 25000 5000 768 10 >> $O/simtopk.log 2>&1
timeout 300 $S simtopk 5000 25000 4608 5 >> $O/simtopk.log 2>&1
cat $O/simtopk.log
for tool in memcheck synccheck; do
  echo "== $tool selftest simtopk 130 300 192 0" >> $O/sanitizer.log
  timeout 300 compute-sanitizer --tool $tool $S simtopk 130 300 192 0 2>&1 | grep -E "ERROR SUMMARY|SIMTOPK" >> $O/sanitizer.log
  echo "== $tool selftest batch 3 1000 768 2 333 256 0" >> $O/sanitizer.log
  timeout 300 compute-sanitizer --tool $tool $S batch 3 1000 768 2 333 256 0 2>&1 | grep -E "ERROR SUMMARY|BATCH" >> $O/sanitizer.log
done
cat $O/sanitizer.log
ncu --set full --clock-control none -k regex:sim_topk -s 1 -c 1 -o $O/simtopk_5000x25000x768 $S simtopk 5000 25000 768 1 > $O/ncu.log 2>&1
timeout 900 $S quick > $O/selftest_quick.log 2>&1; echo "selftest exit $?"; tail -3 $O/selftest_quick.log
