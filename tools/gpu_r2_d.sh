#!/bin/bash
# round 2, GPU call D (1 GPU): ncu captures of the modes, launch list of the bench
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"scan_aggregate|lz4_decode" -c 12 -f -o gpurun_out/r02_modes python tools/profile_modes.py > gpurun_out/d_modes.log 2>&1
ncu -i gpurun_out/r02_modes.ncu-rep --page raw --csv > gpurun_out/r02_modes_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --no-parity > gpurun_out/d_bench_under_ncu.log 2>&1
tail -8 gpurun_out/d_modes.log; wc -l gpurun_out/r02_launches.csv gpurun_out/r02_modes_raw.csv; ls -la gpurun_out | tail -5
