#!/bin/bash
O=/root/repo/gpurun_out/r2q
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
ncu --set full --clock-control none -k regex:syrk_i8x4 -s 1 -c 1 -o $O/syrk_i8x4_36928x3072 $S i8x4 36928 3072 1 1 > $O/ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file $O/launches.csv $S i8x4 36928 3072 0 1 > /dev/null 2>&1
grep -o '"[a-z_0-9:<>, ]*kernel[^"]*","[^"]*","[^"]*","[^"]*","[0-9.]*"$' $O/launches.csv | head -12
cat $O/launches.csv | tail -8 | cut -c1-300
