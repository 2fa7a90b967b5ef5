#!/bin/bash
# round 2, GPU call N: release variants 3 (data dependency), 4 (proxy fence), 5 (one tile late), producer proxy fence
mkdir -p gpurun_out
export SD_TUNE_NO_FRONT_TABLE=1
( for D in "-DSD_EXP_RING=3" "-DSD_EXP_RING=4" "-DSD_EXP_RING=5" "-DSD_EXP_PROD=1" "-DSD_EXP_RING=4 -DSD_EXP_PROD=1" "-DSD_EXP_RING=3 -DSD_EXP_PROD=1"; do
    echo "== defines: [$D]"; SD_JIT_DEFINES="$D" python tools/hash_diag.py 30 2 | grep run
  done ) > gpurun_out/n_hash_diag.txt 2>&1
cat gpurun_out/n_hash_diag.txt
