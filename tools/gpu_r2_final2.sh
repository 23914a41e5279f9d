#!/bin/bash
O=/root/repo/gpurun_out/r2final2
mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-160
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tee $O/gpu_tests.log | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; tail -1 $O/bench_n1.err
python tools/show_bench.py $O/bench_n1.json 2>/dev/null | head -5 | cut -c1-200
