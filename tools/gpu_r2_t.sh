#!/bin/bash
# round 2, GPU call T (1 GPU): final validation of the tree + stored-form e2e with 1 / 2 / 3 copy streams
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/t_pytest.txt
python bench.py > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
for n in 2 3; do SD_TUNE_COPY_STREAMS=$n BENCH_DEBUG=1 python bench.py --steps 5 --warmup 3 --no-cpu --no-also --no-extras --no-parity > gpurun_out/t_e2e_cs$n.json 2> gpurun_out/t_e2e_cs$n.err; done
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/t_smoke.txt 2>&1
tail -4 gpurun_out/t_pytest.txt; tail -2 gpurun_out/t_smoke.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], 'traffic', d['roofline'].get('traffic'), 'e2e', d['e2e']['value'], 'plain', d.get('e2e_plain',{}).get('value'), 'cpu', d['cpu_baseline']['value'], 'parity', d['parity_check']['ok'], 'also', d['also']['value'])
for k in ('also_c4','also_c5'):
    if k in d: print(k, d[k]['value'], d[k]['ms_per_step'], d[k]['roofline']['achieved'], d[k]['parity_check']['ok'])
print('clocks', d['clocks'], 'launches', d['gpu_launches'])
for n in (2,3):
    e=json.loads(open('gpurun_out/t_e2e_cs%d.json'%n).read().strip().splitlines()[-1])
    print('copy streams', n, 'e2e', e['e2e']['value'], e['e2e'].get('ms_per_step'), 'plain', e.get('e2e_plain',{}).get('value'))
PY
