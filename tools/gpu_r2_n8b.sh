#!/bin/bash
# 8 GPUs, final code: the default bench line (reduced legs) for the exchange-step numbers
O=/root/repo/gpurun_out/r2n8b
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 3 --no-variants --no-vitl --no-irtr > $O/bench_n8.json 2> $O/bench_n8.err; tail -1 $O/bench_n8.err
python - <<'PY'
import json
d = json.loads([l for l in open('/root/repo/gpurun_out/r2n8b/bench_n8.json') if l.startswith('{')][0])
print(d['value'], d['ms_per_step'], d['e2e']['value'], 'in-region', d['config']['allreduce_ms_in_timed_region'], 'after barrier', d['config']['allreduce_ms_after_barrier'], d['gram_parity_rel_fro_reduced'], d['merge']['sharded_bit_equal_to_local'], d['merge']['e2e']['value'], d['regmean']['seconds'])
PY
