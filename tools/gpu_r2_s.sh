#!/bin/bash
O=/root/repo/gpurun_out/r2s
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
ncu --set full --clock-control none -k regex:syrk_i8x4 -s 1 -c 1 -o $O/syrk_i8x4_36928x3072 $S i8x4 36928 3072 1 1 > $O/ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file $O/launches.csv $S i8x4 36928 3072 0 1 > /dev/null 2>&1
grep -o 'unnamed>::[a-z0-9_]*' $O/launches.csv | head -0
python - <<'PY'
import csv
rows = list(csv.reader(open('/root/repo/gpurun_out/r2s/launches.csv')))
for r in rows:
    if len(r) > 14 and r[-1].replace('.','').isdigit():
        print(r[4][:60], r[-1])
PY
timeout 900 python -m pytest tests/test_gpu_gram.py -q 2>&1 | tail -2
timeout 900 python bench.py --no-variants --no-vitl --no-irtr --no-cpu-baseline --no-gpu-baseline --no-gramfile --steps 5 > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('/root/repo/gpurun_out/r2s/bench.json') if l.startswith('{')][0])
r = d['regmean']
print({k: r[k] for k in r if k.startswith('e2e_rel') or k.endswith('rel_fro') or k == 'gram_precision_modes'})
PY
for tool in memcheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool $S i8x4 1000 768 0 0 2>&1 | grep -E "ERROR SUMMARY|I8X4" >> $O/sanitizer.log
  timeout 300 compute-sanitizer --tool $tool $S i8x4 333 384 0 1 2>&1 | grep -E "ERROR SUMMARY|I8X4" >> $O/sanitizer.log
done
cat $O/sanitizer.log
