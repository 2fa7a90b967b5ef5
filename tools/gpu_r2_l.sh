#!/bin/bash
# round 2, GPU call L: which change makes the ring + global-table MODE_HASH path exact?  (JIT-time switches, front table off)
mkdir -p gpurun_out
export SD_TUNE_NO_FRONT_TABLE=1
( for D in "" "-DSD_EXP_RING=1" "-DSD_EXP_RING=2" "-DSD_EXP_HASH=1" "-DSD_EXP_RING=1 -DSD_EXP_HASH=1" "-DSD_EXP_RING=2 -DSD_EXP_HASH=1"; do
    echo "== defines: [$D]"
    if [ -z "$D" ]; then python tools/hash_diag.py 30 2 | grep run; else SD_JIT_DEFINES="$D" python tools/hash_diag.py 30 2 | grep run; fi
  done
  echo "== 300 batches, RING=1 HASH=1"; SD_JIT_DEFINES="-DSD_EXP_RING=1 -DSD_EXP_HASH=1" python tools/hash_diag.py 300 2 | grep run
) > gpurun_out/l_hash_diag.txt 2>&1
SD_JIT_DEFINES="-DSD_EXP_RING=2" timeout 500 compute-sanitizer --tool racecheck python tools/hash_diag.py 2 1 > gpurun_out/l_racecheck_ring2.txt 2>&1
SD_JIT_DEFINES="-DSD_EXP_RING=1" timeout 500 compute-sanitizer --tool racecheck python tools/hash_diag.py 2 1 > gpurun_out/l_racecheck_ring1.txt 2>&1
cat gpurun_out/l_hash_diag.txt
for f in ring2 ring1; do echo "-- racecheck $f"; grep -E "run 0|RACECHECK SUMMARY|hazards\]" gpurun_out/l_racecheck_$f.txt | head -8; done
