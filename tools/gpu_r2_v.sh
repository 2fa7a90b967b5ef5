#!/bin/bash
# round 2, GPU call V (1 GPU): coalesced copy window for retained buffers -- parity tests, then stored-form e2e with / without it
mkdir -p gpurun_out
python -m pytest tests/test_gpu_general.py tests/test_gpu_strings.py tests/test_gpu_parity.py tests/test_gpu_edges_and_scale.py -m gpu -q -k "lz4 or compressed or envelope or snappy or resident or submit" 2>&1 | tail -6 > gpurun_out/v_pytest.txt
BENCH_DEBUG=1 python bench.py --steps 5 --warmup 3 --no-also --no-extras --cpu-seconds 8 > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
SD_TUNE_NO_COPY_WINDOW=1 BENCH_DEBUG=1 python bench.py --steps 5 --warmup 3 --no-cpu --no-also --no-extras --no-parity > gpurun_out/v_bench_nowin.json 2> gpurun_out/v_bench_nowin.err
tail -3 gpurun_out/v_pytest.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/v_bench.json').read().strip().splitlines()[-1])
print('window   e2e', d['e2e']['value'], d['e2e'].get('ms_per_step'), 'parity_ok', d['e2e'].get('parity_ok'), 'plain', d.get('e2e_plain',{}).get('value'), 'cpu', d['cpu_baseline']['value'], 'parity', d['parity_check']['ok'], 'value', d['value'])
e=json.loads(open('gpurun_out/v_bench_nowin.json').read().strip().splitlines()[-1])
print('nowindow e2e', e['e2e']['value'], e['e2e'].get('ms_per_step'))
PY
grep "e2e step" gpurun_out/v_bench.err | tail -2; grep "e2e step" gpurun_out/v_bench_nowin.err | tail -2
