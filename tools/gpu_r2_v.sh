#!/bin/bash
# int8x4: TMA fp64 reduce-add epilogue + bit-field digit slicing; ncu of the kernel; API-level tests
O=/root/repo/gpurun_out/r2v
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
for a in "64 128 0 0" "1000 768 0 0" "333 384 0 1" "2560 768 10 0" "2560 3072 10 1" "36928 768 10 0" "36928 3072 10 1" "36928 1024 10 0" "36928 4096 5 0"; do
  timeout 120 $S i8x4 $a 2>&1 | grep -E "I8X4|FAIL|error" | tee -a $O/cases.log
done
ncu --set full --clock-control none --import-source on -k regex:syrk_i8x4 -s 1 -c 1 -o $O/syrk_i8x4_36928x3072 $S i8x4 36928 3072 1 1 > $O/ncu1.log 2>&1
ncu --set full --clock-control none -k regex:syrk_i8x4 -s 1 -c 1 -o $O/syrk_i8x4_36928x768 $S i8x4 36928 768 1 0 > $O/ncu2.log 2>&1
for shape in "36928 3072" "36928 768"; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file $O/launches.csv $S i8x4 $shape 0 1 > /dev/null 2>&1
  python - <<'PY'
import csv
rows = list(csv.reader(open('/root/repo/gpurun_out/r2v/launches.csv')))
for r in rows:
    if len(r) > 14 and r[-1].replace('.','').isdigit():
        print(r[4][:60], r[-1])
PY
done
timeout 900 python -m pytest tests/test_gpu_gram.py tests/test_gpu_regmean_chain.py tests/test_gpu_fused.py -q 2>&1 | tail -3
for tool in memcheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool $S i8x4 1000 768 0 0 2>&1 | grep -E "ERROR SUMMARY|I8X4" >> $O/sanitizer.log
  timeout 300 compute-sanitizer --tool $tool $S i8x4 333 384 0 1 2>&1 | grep -E "ERROR SUMMARY|I8X4" >> $O/sanitizer.log
done
cat $O/sanitizer.log
