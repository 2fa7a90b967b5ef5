#!/bin/bash
# round 2, GPU call U (1 GPU): sanitizer passes over the code added in round 2 (final tree)
mkdir -p gpurun_out
K="written_on_the_device or projection or grows_between or encoder or raw_string or min_max or decimal or casts or null_key or string_key"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_general.py tests/test_gpu_strings.py tests/test_gpu_encoder.py tests/test_gpu_known_answers.py -m gpu -q -k "$K" > gpurun_out/u_memcheck.txt 2>&1
timeout 700 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_general.py tests/test_gpu_strings.py tests/test_gpu_encoder.py -m gpu -q -k "written_on_the_device or projection_of_raw or grows_between or encoder or update_deltas" > gpurun_out/u_racecheck.txt 2>&1
timeout 300 compute-sanitizer --tool initcheck python -m pytest tests/test_gpu_general.py -m gpu -q -k "written_on_the_device or projection_with_filter" > gpurun_out/u_initcheck.txt 2>&1
for f in memcheck racecheck initcheck; do echo "-- $f"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|Uninit|hazard" gpurun_out/u_$f.txt | head -8; done
