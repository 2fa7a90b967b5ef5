#!/bin/bash
# round 2, GPU call J: hash-mode diagnostic at scale (every group against numpy), GPU tests after the ensure_out fix
mkdir -p gpurun_out
( SD_TUNE_NO_FRONT_TABLE=1 python tools/hash_diag.py 300 3
  SD_TUNE_NO_FRONT_TABLE=1 SD_TUNE_NSTAGES=2 python tools/hash_diag.py 300 2
  python tools/hash_diag.py 300 2
  SD_TUNE_NO_FRONT_TABLE=1 python tools/hash_diag.py 30 3 ) > gpurun_out/j_hash_diag.txt 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/j_pytest.txt
cat gpurun_out/j_hash_diag.txt; tail -8 gpurun_out/j_pytest.txt
