#!/bin/bash
O=/root/repo/gpurun_out/r2j
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
ncu --set full --clock-control none --import-source on -k regex:syrk_f64 -s 2 -c 1 -o $O/syrk_f64_36928x768 $S f64 f32 36928 768 1 0 > $O/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:regmean_rhs_pipelined -s 1 -c 1 -o $O/rhs_768x3072 $S rhs 768 3072 3 > $O/ncu2.log 2>&1
ls -la $O
