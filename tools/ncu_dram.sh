#!/bin/bash
# usage: tools/ncu_dram.sh <tag> <selftest args...>   -> prints duration + DRAM bytes of the 2nd syrk_2sm launch
tag=$1; shift
ncu --set full --clock-control none -k regex:syrk_2sm -s 1 -c 1 -f -o gpurun_out/$tag vl-merging_b200/csrc/build/selftest "$@" > /dev/null 2>&1
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr,units,vals=rows[0],rows[1],rows[2]
out=[]
for w in ('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct'):
    i=hdr.index(w); out.append(f'{w.split(\".\")[0]}={vals[i]}{units[i]}')
print('$tag', ' '.join(out))
"
