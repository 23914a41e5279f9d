#!/bin/bash
O=/root/repo/gpurun_out/r2n
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_merge.py tests/test_gpu_random_structures.py tests/test_gpu_large.py tests/test_gpu_e2e.py -q 2>&1 | tail -4
timeout 600 python bench.py --no-variants --no-vitl --no-irtr --no-regmean --no-cpu-baseline --no-gpu-baseline --steps 3 > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('/root/repo/gpurun_out/r2n/bench.json') if l.startswith('{')][0])
print(d['merge']['value'], d['merge']['e2e'])
PY
