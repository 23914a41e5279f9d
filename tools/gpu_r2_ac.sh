#!/bin/bash
# 2 GPUs: batched pack/unpack in the exchange step (NCCL parity test), gramfile tests, exchange time
O=/root/repo/gpurun_out/r2ac
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_gramfile.py tests/test_gpu_gram.py -q 2>&1 | tee $O/pytest.log | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-variants --no-vitl --no-irtr --no-regmean > $O/bench_n2.json 2> $O/bench_n2.err; tail -2 $O/bench_n2.err
python - <<'PY'
import json
d = json.loads([l for l in open('/root/repo/gpurun_out/r2ac/bench_n2.json') if l.startswith('{')][0])
print(d['value'], d['config']['allreduce_ms_in_timed_region'], d['config']['allreduce_ms_after_barrier'], d['gram_parity_rel_fro_reduced'])
PY
