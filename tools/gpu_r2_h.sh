#!/bin/bash
# round 2, GPU call H: LZ4 wide copies (correctness + throughput), C4 / C5 after the lock / timing fixes
mkdir -p gpurun_out
python -m pytest tests/test_gpu_general.py tests/test_gpu_strings.py tests/test_gpu_encoder.py -m gpu -q -k "lz4 or compressed or errors or encoder" 2>&1 | tail -15 > gpurun_out/h_pytest.txt
( LZ4_KINDS=0,1,2,3,4 LZ4_VARIANTS=3 python tools/lz4_bench.py 3000 200000 3; LZ4_KINDS=0 LZ4_VARIANTS=3 python tools/lz4_bench.py 6000 200000 3 ) > gpurun_out/h_lz4.txt 2>&1
SD_DEBUG_TIMING=1 python bench.py --workload c4 --steps 3 --warmup 1 > gpurun_out/h_c4.json 2> gpurun_out/h_c4.err
python bench.py --workload c5 --steps 20 --warmup 3 > gpurun_out/h_c5.json 2> gpurun_out/h_c5.err
BENCH_DEBUG=1 python bench.py --steps 10 --warmup 3 --no-cpu --no-also --no-parity --no-extras > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_general.py -m gpu -q -k hard_blocks > gpurun_out/h_racecheck_lz4.txt 2>&1
tail -5 gpurun_out/h_pytest.txt; cat gpurun_out/h_lz4.txt; grep finish_project gpurun_out/h_c4.err | tail -6; grep "e2e step" gpurun_out/h_bench.err | tail -3; tail -4 gpurun_out/h_racecheck_lz4.txt
python - <<'PY'
import json
for f in ('h_c4','h_c5'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'], d['parity_check']['ok'])
    except Exception as e: print(f, 'ERR', e)
d=json.loads(open('gpurun_out/h_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step']); 
for k in ('e2e','e2e_plain'): print(k, d[k]['value'], d[k]['ms_per_step'])
PY
