#!/bin/bash
O=/root/repo/gpurun_out/r2r
mkdir -p $O
bash tools/gpu_r2_p.sh | grep "I8X4\|EXIT"
timeout 900 python -m pytest tests/test_gpu_gram.py tests/test_gpu_regmean_chain.py -q -s > $O/log.txt 2>&1
grep -E "^E  |int8x4 alpha|passed|failed" $O/log.txt | cut -c1-330 | head -20
