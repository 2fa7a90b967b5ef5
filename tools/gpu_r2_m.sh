#!/bin/bash
# round 2, GPU call M: is the STAGED data wrong?  (verify mode: every staged value against global memory, hash workload)
mkdir -p gpurun_out
export SD_DEBUG_VERIFY=1 SD_JIT_DEFINES="-DSD_EXP_VERIFY=1"
( echo "== front off"; SD_TUNE_NO_FRONT_TABLE=1 python tools/hash_diag.py 30 3 2>&1 | grep -E "run|verify"
  echo "== front on";  python tools/hash_diag.py 30 2 2>&1 | grep -E "run|verify"
  echo "== front off, 2 stages"; SD_TUNE_NO_FRONT_TABLE=1 SD_TUNE_NSTAGES=2 python tools/hash_diag.py 30 2 2>&1 | grep -E "run|verify"
) > gpurun_out/m_verify.txt 2>&1
cat gpurun_out/m_verify.txt
