#!/bin/bash
# round 2, GPU call G (1 GPU): whole GPU suite after the barrier-free NULL / overlay paths, racecheck, C4 / C5 workloads, ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/g_pytest.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/g_racecheck.txt 2>&1
python bench.py --workload c4 --steps 5 --warmup 2 > gpurun_out/g_c4.json 2> gpurun_out/g_c4.err
python bench.py --workload c5 --steps 20 --warmup 3 > gpurun_out/g_c5.json 2> gpurun_out/g_c5.err
ncu --set full --clock-control none --import-source on -k regex:"scan_aggregate|lz4_decode" -c 12 -f -o gpurun_out/r02_modes python tools/profile_modes.py > gpurun_out/g_modes.log 2>&1
ncu -i gpurun_out/r02_modes.ncu-rep --page raw --csv > gpurun_out/r02_modes_raw.csv 2>/dev/null
tail -6 gpurun_out/g_pytest.txt; tail -3 gpurun_out/g_racecheck.txt; tail -2 gpurun_out/g_c4.err gpurun_out/g_c5.err
python - <<'PY'
import json
for f in ('g_c4','g_c5'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'], d['parity_check']['ok'])
    except Exception as e: print(f, 'ERR', e)
PY
grep -v "^==PROF==" gpurun_out/g_modes.log | tail -8; ls -la gpurun_out/r02_modes*
