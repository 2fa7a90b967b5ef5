#!/bin/bash
# round 2, GPU call S (1 GPU): C5 A/B (incremental scan on / off), full ncu capture of the Q1 kernel
mkdir -p gpurun_out
python bench.py --workload c5 --steps 20 --warmup 3 > gpurun_out/s_c5_incr.json 2> gpurun_out/s_c5_incr.err
SD_TUNE_NO_INCREMENTAL_SCAN=1 python bench.py --workload c5 --steps 20 --warmup 3 > gpurun_out/s_c5_full.json 2> gpurun_out/s_c5_full.err
python tools/profile_modes.py 24 2>&1 | grep -E "c5 overlay|c4 project|hash group-by" > gpurun_out/s_modes.txt
ncu --set full --clock-control none --import-source on -k regex:"scan_aggregate" -s 1 -c 1 -f -o gpurun_out/r02_q1 python bench.py --steps 1 --warmup 1 --rows 200000000 --no-cpu --no-extras --no-parity --no-e2e --no-also > gpurun_out/s_q1_ncu.log 2>&1
ncu -i gpurun_out/r02_q1.ncu-rep --page raw --csv > gpurun_out/r02_q1_raw.csv 2>/dev/null
python - <<'PY'
import json
for f in ('s_c5_incr','s_c5_full'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['kernel_ms_per_launch'], d['parity_check']['ok'], d.get('ingested_batches_during_timed_region'), d.get('snapshots_batches'))
PY
cat gpurun_out/s_modes.txt; ls -la gpurun_out/r02_q1.ncu-rep; wc -l gpurun_out/r02_q1_raw.csv
