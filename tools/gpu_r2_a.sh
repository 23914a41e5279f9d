#!/bin/bash
# round-2 GPU check A: the cta_group::2 SYRK against the multicast pair kernel, and the split mode
cd vl-merging_b200/csrc
S=build/selftest
O=/root/repo/gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
for v in 3 2; do
  for shape in "f32 36928 3072 10 1" "f32 36928 768 20 0" "bf16 36928 3072 10 1" "f32 2560 3072 20 1" "f32 36928 4096 10 1" "f32 1000 768 0 0"; do
    echo "== variant $v case $shape" >> $O/cases.log
    VLM_SYRK_VARIANT=$v timeout 120 $S case $shape >> $O/cases.log 2>&1 || echo "EXIT $?" >> $O/cases.log
  done
done
echo "== split" >> $O/cases.log
for shape in "1000 768 0 0" "36928 768 20 0" "36928 3072 10 1" "2560 3072 20 1"; do
  timeout 120 $S split $shape >> $O/cases.log 2>&1 || echo "EXIT $?" >> $O/cases.log
done
timeout 600 $S quick > $O/selftest_quick.log 2>&1; echo "selftest exit $?" >> $O/cases.log
cat $O/cases.log
tail -5 $O/selftest_quick.log
