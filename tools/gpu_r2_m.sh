#!/bin/bash
O=/root/repo/gpurun_out/r2m
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_gram.py -q 2>&1 | tail -15
