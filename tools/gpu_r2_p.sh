#!/bin/bash
O=/root/repo/gpurun_out/r2p
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
for shape in "64 128 0 0" "1000 768 0 0" "333 384 0 1" "2560 768 10 0" "2560 3072 10 1" "36928 768 10 0" "36928 3072 5 1"; do
  timeout 120 $S i8x4 $shape >> $O/i8.log 2>&1 || echo "EXIT $? for $shape" >> $O/i8.log
done
cat $O/i8.log
