#!/bin/bash
O=/root/repo/gpurun_out/r2e
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "gpu tests exit $?" >> $O/gpu_tests.log
tail -40 $O/gpu_tests.log
python tools/time_regmean.py base > $O/regmean_streams.log 2>&1
cat $O/regmean_streams.log
