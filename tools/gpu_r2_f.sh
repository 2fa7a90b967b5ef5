#!/bin/bash
# round 2, GPU call F (N GPUs): the driver's multi-rank launch of bench.py
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/f_bench_n$N.json 2> gpurun_out/f_bench_n$N.err
tail -5 gpurun_out/f_bench_n$N.err; python - <<PY
import json
d=json.loads(open('gpurun_out/f_bench_n$N.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','scaling','n_gpus'): print(k, d[k])
for k in ('e2e','e2e_plain','e2e_pageable_unretained','cpu_baseline','exchange'):
    if k in d: print(k, {a:b for a,b in d[k].items() if a in ('value','ms_per_step','h2d_bytes_per_step','parity_ok','cores','world','slot_bytes','all_gathers','regrows')})
print('parity', {a:b for a,b in d.get('parity_check',{}).items() if a in ('ok','rows','max_rel_err','counts_exact','resident_store_result')})
print('also', d['also']['value'], d['also']['ms_per_step'], d['also'].get('parity_check',{}).get('ok'))
for k in ('also_c4','also_c5'):
    if k in d: print(k, {a:b for a,b in d[k].items() if a in ('value','ms_per_step','rows_out')}, d[k]['roofline']['achieved'], d[k]['parity_check']['ok'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'])
PY
