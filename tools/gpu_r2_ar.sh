#!/bin/bash
# N = 8: exchange step, default (pack + NCCL + unpack) vs one multimem kernel over symmetric arenas
N=$(nvidia-smi -L | wc -l)
O=/root/repo/gpurun_out/r2ar
mkdir -p $O
for mode in "" "--symmetric"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tools/exchange_trace.py $mode 2>&1 | grep -v "OMP\|\*\*\*" | tee -a $O/trace.log | tail -8
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 10 --warmup 3 --no-variants --no-vitl --no-irtr --no-regmean --symmetric > $O/bench_n${N}_symmetric.json 2> $O/bench.err; tail -2 $O/bench.err | grep -v OMP
python - "$O/bench_n${N}_symmetric.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][0])
    print(d['value'], 'in-region', d['config']['allreduce_ms_in_timed_region'], 'after barrier', d['config']['allreduce_ms_after_barrier'], 'parity', d['gram_parity_rel_fro_reduced'], '|', d['config']['allreduce'][:60])
except Exception as e:
    print('no line', e)
PY
