#!/bin/bash
# prints, for the fp32 single-problem SYRK kernel, the number of SASS instructions between consecutive UTCHMMA
# (the single-thread MMA issue loop is instruction-issue bound: fewer is better)
obj=${1:-vl-merging_b200/csrc/build/syrk_pair.o}
cuobjdump -sass $obj | awk '/Function :/{f=($0 ~ /ILi4ELi2ELb0E/)} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | awk '
  {n++} /UTCHMMA|UTCMMA/{ if (last) printf "%d ", n-last; last=n } /UTCBAR/{ if (last) {printf "| to commit %d\n", n-last; last=0} }'
echo
