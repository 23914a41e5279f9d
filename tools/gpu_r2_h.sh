#!/bin/bash
O=/root/repo/gpurun_out/r2h
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
timeout 600 python -m pytest tests/test_gpu_simtopk.py -q > $O/simtopk.log 2>&1; echo "exit $?" >> $O/simtopk.log
tail -5 $O/simtopk.log
timeout 600 python bench.py --no-variants --no-vitl --no-cpu-baseline --no-gpu-baseline --no-gramfile --no-regmean --steps 5 > $O/bench_irtr.json 2> $O/bench_irtr.err; echo "bench exit $?" >> $O/bench_irtr.err
tail -3 $O/bench_irtr.err
python - <<'PY'
import json
d = json.loads([l for l in open('/root/repo/gpurun_out/r2h/bench_irtr.json') if l.startswith('{')][0])
print(d['irtr'])
PY
# compute-sanitizer on the new kernels (small shapes)
for tool in memcheck synccheck; do
  for args in "case f32 1000 768 0 0" "case bf16 333 256 0 1" "split 1000 768 0 0" "f64 f32 1000 768 0 0" "strided f32 4 617 40 577 768 0"; do
    echo "== $tool selftest $args" >> $O/sanitizer.log
    timeout 300 compute-sanitizer --tool $tool $S $args 2>&1 | grep -E "ERROR SUMMARY|Error|error|OK|FAIL" | head -8 >> $O/sanitizer.log
  done
done
cat > /tmp/st.py <<'PY'
import torch, vl_merging_b200 as vlm
a = torch.randn(130, 192, device="cuda").half(); b = torch.randn(300, 192, device="cuda").half()
v, i = vlm.sim_topk(a, b, 10); torch.cuda.synchronize()
w = (a.float() @ b.float().t()).topk(10, dim=1)
print("simtopk ok", bool((i == w.indices).float().mean() > 0.99))
PY
echo "== memcheck sim_topk" >> $O/sanitizer.log
timeout 600 compute-sanitizer --tool memcheck python /tmp/st.py 2>&1 | grep -E "ERROR SUMMARY|simtopk" >> $O/sanitizer.log
cat $O/sanitizer.log
# ncu launch list of one calibration step
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --profile --steps 1 --warmup 2 > $O/ncu_launch.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launches_summary.txt 2>&1; head -30 $O/launches_summary.txt
