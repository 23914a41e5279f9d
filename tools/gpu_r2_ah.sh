#!/bin/bash
O=/root/repo/gpurun_out/r2ah
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gramfile.py tests/test_gpu_fused.py tests/test_gpu_e2e.py -q 2>&1 | tail -8
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
