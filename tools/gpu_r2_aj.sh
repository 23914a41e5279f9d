#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_large.py -q 2>&1 | tail -3
