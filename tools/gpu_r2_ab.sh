#!/bin/bash
# the headline in the RegMean-grade int8x4 mode on one GPU
O=/root/repo/gpurun_out/r2ab
mkdir -p $O
timeout 1200 python bench.py --gram-precision int8x4 --no-variants --no-vitl --no-irtr --no-gpu-baseline --steps 10 > $O/bench_n1_int8x4.json 2> $O/bench_n1_int8x4.err; tail -3 $O/bench_n1_int8x4.err
python - <<'PY'
import json
d = json.loads([l for l in open('/root/repo/gpurun_out/r2ab/bench_n1_int8x4.json') if l.startswith('{')][0])
for k in ('value','ms_per_step','dtype','e2e','gpu_launches','gram_parity_rel_fro'): print(k, d[k])
r = d['roofline']; print({k: r[k] for k in r if k != 'kernel'})
print({k: d['regmean'][k] for k in ('seconds','e2e_rel_err_vs_fp64_grams_int8x4','check_rel_err_vs_torch_fp64')})
PY
