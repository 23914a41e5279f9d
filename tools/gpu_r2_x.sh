#!/bin/bash
# what limits the grouped (batched) SYRK launches: ncu of the image group (9 x 36928x768) and the text group
O=/root/repo/gpurun_out/r2x
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
ncu --set full --clock-control none --import-source on -k regex:syrk_2sm -s 10 -c 1 -o $O/batch_image9 $S batch 9 36928 768 0 0 0 1 > $O/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:syrk_2sm -s 49 -c 1 -o $O/batch_text48 $S batch 36 2560 768 12 2560 3072 1 > $O/ncu2.log 2>&1
ls -la $O
