#!/bin/bash
# round 2, GPU call R (1 GPU): overlay windows (tests, C5), racecheck with per-thread release, profiles of the shipped kernels
mkdir -p gpurun_out
python -m pytest tests/test_gpu_general.py tests/test_gpu_known_answers.py tests/test_gpu_parity.py -m gpu -q -k "delta or delete or update or basic_delete or stats or hybrid or grows" 2>&1 | tail -8 > gpurun_out/r_pytest_overlay.txt
python bench.py --workload c5 --steps 20 --warmup 3 > gpurun_out/r_c5.json 2> gpurun_out/r_c5.err
timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_general.py -m gpu -q -k "update_deltas_and_delete_mask" > gpurun_out/r_racecheck_overlay.txt 2>&1
SD_JIT_DEFINES="-DSD_EXP_ARRIVE_ALL" timeout 400 compute-sanitizer --tool racecheck python tools/hash_diag.py 2 1 > gpurun_out/r_racecheck_arrive_all.txt 2>&1
SD_JIT_DEFINES="-DSD_EXP_ARRIVE_ALL" python tools/hash_diag.py 300 2 > gpurun_out/r_arrive_all_diag.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --no-parity --no-e2e > gpurun_out/r_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"scan_aggregate" -s 3 -c 1 -f -o gpurun_out/r02_q1 python bench.py --steps 1 --warmup 1 --rows 200000000 --no-cpu --no-extras --no-parity --no-e2e --no-also > gpurun_out/r_q1_ncu.log 2>&1
ncu -i gpurun_out/r02_q1.ncu-rep --page raw --csv > gpurun_out/r02_q1_raw.csv 2>/dev/null
tail -4 gpurun_out/r_pytest_overlay.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r_c5.json').read().strip().splitlines()[-1])
print('c5', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'], d['parity_check']['ok'])
PY
for f in overlay arrive_all; do echo "-- racecheck $f"; grep -E "passed|failed| run |RACECHECK SUMMARY|hazards\]" gpurun_out/r_racecheck_$f.txt | head -6; done
cat gpurun_out/r_arrive_all_diag.txt | grep " run "
wc -l gpurun_out/r02_launches.csv gpurun_out/r02_q1_raw.csv; ls -la gpurun_out/r02_q1.ncu-rep
