#!/bin/bash
# round 2, GPU call C: whole GPU suite (strings, decimals, casts), LZ4 after the dependency-mask rounds, default bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/c_pytest.txt
( LZ4_KINDS=0,1,2,3,4 LZ4_VARIANTS=3 python tools/lz4_bench.py 3000 200000 3; LZ4_KINDS=0 LZ4_VARIANTS=3 python tools/lz4_bench.py 6000 200000 3 ) > gpurun_out/c_lz4.txt 2>&1
BENCH_DEBUG=1 python bench.py --steps 20 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
tail -30 gpurun_out/c_pytest.txt; cat gpurun_out/c_lz4.txt; grep "e2e step" gpurun_out/c_bench.err | tail -8; tail -3 gpurun_out/c_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step'): print(k, d[k])
for k in ('e2e','e2e_plain','e2e_pageable_unretained','cpu_baseline'):
    if k in d: print(k, {a:b for a,b in d[k].items() if a in ('value','ms_per_step','h2d_bytes_per_step','parity_ok','cores')})
print('parity', {a:b for a,b in d.get('parity_check',{}).items() if a in ('ok','rows','max_rel_err','counts_exact')})
print('also', d['also']['value'], d['also']['ms_per_step'], d['also'].get('parity_check',{}).get('ok'))
print('roofline', d['roofline']['achieved'], d['roofline']['frac'])
PY
