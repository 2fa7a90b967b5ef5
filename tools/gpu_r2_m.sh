#!/bin/bash
# round 2, GPU call M3: shape of the damage (per warp / per part of the tile / per column)
mkdir -p gpurun_out
export SD_DEBUG_VERIFY=1 SD_JIT_DEFINES="-DSD_EXP_VERIFY=1"
( echo "== front off"; SD_TUNE_NO_FRONT_TABLE=1 python tools/hash_diag.py 30 2 2>&1 | grep -E "verify"
  echo "== front off, 2 stages"; SD_TUNE_NO_FRONT_TABLE=1 SD_TUNE_NSTAGES=2 python tools/hash_diag.py 30 1 2>&1 | grep -E "verify"
  echo "== front off, 3 batches"; SD_TUNE_NO_FRONT_TABLE=1 python tools/hash_diag.py 3 1 2>&1 | grep -E "verify"
) > gpurun_out/m_verify3.txt 2>&1
cat gpurun_out/m_verify3.txt
