#!/bin/bash
# round 2, GPU call E: new tests (encoder, wide keys), default bench with the C4 / C5 extras
mkdir -p gpurun_out
python -m pytest tests/test_gpu_known_answers.py tests/test_gpu_encoder.py tests/test_gpu_strings.py -m gpu -q 2>&1 | tail -40 > gpurun_out/e_pytest.txt
BENCH_DEBUG=1 python bench.py --steps 20 --warmup 3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
tail -25 gpurun_out/e_pytest.txt; grep "e2e step" gpurun_out/e_bench.err | tail -6; grep -v "e2e step\|^\[rank" gpurun_out/e_bench.err | tail -8; python - <<'PY'
import json
d=json.loads(open('gpurun_out/e_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step'): print(k, d[k])
for k in ('e2e','e2e_plain','e2e_pageable_unretained','cpu_baseline'):
    if k in d: print(k, {a:b for a,b in d[k].items() if a in ('value','ms_per_step','h2d_bytes_per_step','parity_ok','cores')})
print('parity', {a:b for a,b in d.get('parity_check',{}).items() if a in ('ok','rows','max_rel_err','counts_exact')})
print('also', d['also']['value'], d['also']['ms_per_step'], d['also'].get('parity_check',{}).get('ok'))
for k in ('also_c4','also_c5'):
    if k in d: print(k, {a:b for a,b in d[k].items() if a in ('value','ms_per_step','rows_out','selectivity','snapshots_batches','ingested_batches_during_timed_region')}, d[k]['roofline']['achieved'], d[k]['roofline']['frac'], d[k]['parity_check'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'])
PY
