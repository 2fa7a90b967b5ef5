#!/bin/bash
# round 2, GPU call A: tests, probe of the new execution path, LZ4 variants, racecheck, full bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/a_pytest.txt
python tools/r2_probe.py both 30 > gpurun_out/a_probe.txt 2>&1
python tools/lz4_bench.py 1024 200000 3 > gpurun_out/a_lz4.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/a_racecheck.txt 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
tail -5 gpurun_out/a_pytest.txt; cat gpurun_out/a_probe.txt; tail -30 gpurun_out/a_lz4.txt; tail -3 gpurun_out/a_racecheck.txt; tail -c 3000 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
