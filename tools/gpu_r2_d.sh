#!/bin/bash
# round-2 GPU check D: full GPU suite, default bench line
O=/root/repo/gpurun_out/r2d
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/gpu_tests.log 2>&1; echo "gpu tests exit $?" >> $O/gpu_tests.log
tail -30 $O/gpu_tests.log
( time timeout 900 python bench.py ) > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench exit $?" >> $O/bench_n1.err
tail -5 $O/bench_n1.err
python tools/show_bench.py $O/bench_n1.json 2>/dev/null | head -80 || head -c 3000 $O/bench_n1.json
