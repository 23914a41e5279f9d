#!/bin/bash
O=/root/repo/gpurun_out/r2al
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dist.py -q -x 2>&1 | tee $O/pytest.log | tail -15
