#!/bin/bash
O=/root/repo/gpurun_out/r2i
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
for f in 0 1 2 4; do
  echo "== proportional factor $f" >> $O/batch.log
  VLM_SYRK_BATCH_PROPORTIONAL=$f timeout 200 $S batch 9 36928 768 0 0 0 10 >> $O/batch.log 2>&1
  VLM_SYRK_BATCH_PROPORTIONAL=$f timeout 200 $S batch 36 2560 768 12 2560 3072 10 >> $O/batch.log 2>&1
  VLM_SYRK_BATCH_PROPORTIONAL=$f timeout 200 $S batch 36 2560 768 0 0 0 10 >> $O/batch.log 2>&1
  VLM_SYRK_BATCH_PROPORTIONAL=$f timeout 200 $S batch 3 1000 768 2 333 256 0 >> $O/batch.log 2>&1
done
cat $O/batch.log
