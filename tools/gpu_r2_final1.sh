#!/bin/bash
# final round-2 evidence on one B200: smoke, GPU test suite, default bench line + reference arm, launch list, sanitizer
O=/root/repo/gpurun_out/r2final
mkdir -p $O
S=vl-merging_b200/csrc/build/selftest
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.log
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tee $O/gpu_tests.log | tail -3
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; tail -1 $O/bench_reference_arm.json | cut -c1-300
timeout 1800 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; tail -2 $O/bench_n1.err
python tools/show_bench.py $O/bench_n1.json 2>/dev/null | head -12 | cut -c1-400
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --profile --steps 1 --warmup 2 > $O/ncu_launch.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launches_summary.txt 2>&1; head -12 $O/launches_summary.txt
for tool in memcheck synccheck; do
  for args in "i8x4 1000 768 0 0" "i8x4 333 384 0 1 f16" "case f32 1000 768 0 0" "f64 f32 1000 768 0 0" "pack"; do
    echo "== $tool selftest $args" >> $O/sanitizer.log
    timeout 300 compute-sanitizer --tool $tool $S $args 2>&1 | grep -E "ERROR SUMMARY" | head -3 >> $O/sanitizer.log
  done
done
grep -c "0 errors" $O/sanitizer.log; grep -v "0 errors" $O/sanitizer.log | grep ERROR
