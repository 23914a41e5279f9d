#!/bin/bash
# int8x4 mode: does a side stream for the Gram launches help (SYRK share is 41 % there)?
O=/root/repo/gpurun_out/r2ad
mkdir -p $O
timeout 1200 python bench.py --gram-precision int8x4 --no-vitl --no-irtr --no-gpu-baseline --no-cpu-baseline --no-regmean --steps 10 > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('/root/repo/gpurun_out/r2ad/bench.json') if l.startswith('{')][0])
print(d['value'], d['ms_per_step'])
for k, v in d['forward_variants'].items(): print(k, {a: v[a] for a in v if a in ('value', 'ms_per_step', 'syrk_tflops')})
print(d['roofline']['by_shape'])
PY
