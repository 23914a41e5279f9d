#!/bin/bash
# round-2 full check: GPU suite, smoke, both bench arms as the driver runs them
O=/root/repo/gpurun_out/r2l
mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/gpu_tests.log 2>&1; echo "gpu tests exit $?" >> $O/gpu_tests.log
tail -12 $O/gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -2 $O/smoke.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 ) > $O/bench_ref.json 2> $O/bench_ref.err; tail -4 $O/bench_ref.err
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 ) > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench exit $?" >> $O/bench_n1.err
tail -5 $O/bench_n1.err
python tools/show_bench.py $O/bench_ref.json $O/bench_n1.json | cut -c1-900
