#!/bin/bash
# RegMean in the difference form: selftest, merge / chain / e2e tests, bench regmean numbers
O=/root/repo/gpurun_out/r2aa
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
timeout 600 $S quick > $O/selftest_quick.log 2>&1; grep -E "REGMEAN|FAIL|launches=" $O/selftest_quick.log
timeout 900 python -m pytest tests/test_gpu_merge.py tests/test_gpu_regmean_chain.py tests/test_gpu_e2e.py tests/test_gpu_large.py tests/test_gpu_fused.py tests/test_gpu_gramfile.py -q 2>&1 | tail -3
timeout 900 python bench.py --no-variants --no-irtr --no-cpu-baseline --no-gpu-baseline --no-gramfile --steps 5 > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('/root/repo/gpurun_out/r2aa/bench.json') if l.startswith('{')][0])
r = d['regmean']
print({k: r[k] for k in r if k not in ('e2e_detail', 'note', 'e2e_note')})
print(d['vitl'])
PY
