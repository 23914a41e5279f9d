#!/bin/bash
# 8 GPUs: the default bench line as the driver launches it, then the RegMean-grade int8x4 mode
O=/root/repo/gpurun_out/r2n8
mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; tail -2 $O/bench_n8.err
python tools/show_bench.py $O/bench_n8.json 2>/dev/null | head -6 | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --gram-precision int8x4 --no-variants --no-vitl --no-irtr > $O/bench_n8_int8x4.json 2> $O/bench_n8_int8x4.err; tail -2 $O/bench_n8_int8x4.err
python - <<'PY'
import json
for f in ('bench_n8.json', 'bench_n8_int8x4.json'):
    try:
        d = json.loads([l for l in open('/root/repo/gpurun_out/r2n8/' + f) if l.startswith('{')][0])
    except Exception as e:
        print(f, 'no line', e); continue
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['allreduce_ms_in_timed_region'], d['config']['allreduce_ms_after_barrier'], d['config']['allreduce'], d['gram_parity_rel_fro_reduced'], d['merge']['sharded_bit_equal_to_local'], d['regmean']['seconds'])
PY
