#!/bin/bash
O=/root/repo/gpurun_out/r2as
mkdir -p $O
python tools/packbench.py | tee $O/packbench.log
ncu --set full --clock-control none -k regex:sym_ -s 6 -c 3 -o $O/packkernels python tools/packbench.py > $O/ncu.log 2>&1
ls -la $O
