#!/bin/bash
# round 2, GPU call I: MODE_HASH front table, device row writer (C4), LZ4 back to the dependency-mask kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/i_pytest.txt
timeout 600 python tools/hash_bench.py 300 > gpurun_out/i_hash.txt 2>&1
python bench.py --workload c4 --steps 5 --warmup 2 > gpurun_out/i_c4.json 2> gpurun_out/i_c4.err
SD_TUNE_HOST_ROWS=1 python bench.py --workload c4 --steps 3 --warmup 1 > gpurun_out/i_c4_host.json 2> gpurun_out/i_c4_host.err
python bench.py --workload c5 --steps 20 --warmup 3 > gpurun_out/i_c5.json 2> gpurun_out/i_c5.err
( LZ4_KINDS=0 LZ4_VARIANTS=3 python tools/lz4_bench.py 3000 200000 3; LZ4_KINDS=0 LZ4_VARIANTS=3 python tools/lz4_bench.py 6000 200000 3 ) > gpurun_out/i_lz4.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_general.py -m gpu -q -k "front_table or written_on_the_device or hard_blocks" > gpurun_out/i_racecheck.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_general.py tests/test_gpu_strings.py -m gpu -q -k "front_table or written_on_the_device or projection_of_raw" > gpurun_out/i_memcheck.txt 2>&1
tail -5 gpurun_out/i_pytest.txt; cat gpurun_out/i_hash.txt; cat gpurun_out/i_lz4.txt; tail -4 gpurun_out/i_racecheck.txt; tail -4 gpurun_out/i_memcheck.txt
python - <<'PY'
import json
for f in ('i_c4','i_c4_host','i_c5'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'], d['parity_check']['ok'], d.get('d2h_bytes_per_step'))
    except Exception as e: print(f, 'ERR', e)
PY
