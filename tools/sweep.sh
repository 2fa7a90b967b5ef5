#!/bin/bash
# tile-shape sweep of the scan kernel through the NVRTC path (SD_TUNE_* change the plan signature)
for q in q1 q6; do
  for rpt in 2 4 8; do
    for mc in 1 2 3 4; do
      echo -n "$q rpt=$rpt minctas=$mc : "
      SD_TUNE_RPT=$rpt SD_TUNE_MIN_CTAS=$mc timeout 120 python tools/profile_scan.py $q ${1:-200000000} 4 2>&1 | tail -1
    done
  done
done
