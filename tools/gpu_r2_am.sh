#!/bin/bash
# the exchange as one multimem kernel (symmetric arenas) vs the NCCL path, N = number of GPUs on the box
O=/root/repo/gpurun_out/r2am
mkdir -p $O
N=$(nvidia-smi -L | wc -l)
for mode in "--symmetric" ""; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 10 --warmup 3 --no-variants --no-vitl --no-irtr --no-regmean $mode > $O/bench_n${N}${mode}.json 2> $O/bench_n${N}${mode}.err; tail -2 $O/bench_n${N}${mode}.err | grep -v OMP
  python - "$O/bench_n${N}${mode}.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][0])
    print(d['value'], 'in-region', d['config']['allreduce_ms_in_timed_region'], 'after barrier', d['config']['allreduce_ms_after_barrier'], 'parity', d['gram_parity_rel_fro_reduced'], '|', d['config']['allreduce'][:90])
except Exception as e:
    print('no line', e)
PY
done
