#!/bin/bash
# int8x4 epilogue v2 (one fp64 add per element per phase, exponent-bit scaling) + pre-pass with 4 loads in flight
O=/root/repo/gpurun_out/r2t
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
for epi in 1 0; do export VLM_I8_TMA_EPILOGUE=$epi; echo "TMA epilogue: $epi" | tee -a $O/cases.log
for a in "64 128 0 0" "1000 768 0 0" "333 384 0 1" "2560 768 10 0" "2560 3072 10 1" "36928 768 10 0" "36928 3072 10 1" "36928 1024 10 0" "36928 4096 5 0"; do
  timeout 300 $S i8x4 $a 2>&1 | grep -E "I8X4|FAIL|error" | tee -a $O/cases.log
done
for shape in "36928 3072" "36928 768"; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file $O/launches.csv $S i8x4 $shape 0 1 > /dev/null 2>&1
  python - <<'PY'
import csv
rows = list(csv.reader(open('/root/repo/gpurun_out/r2t/launches.csv')))
for r in rows:
    if len(r) > 14 and r[-1].replace('.','').isdigit():
        print(r[4][:60], r[-1])
PY
done
