#!/bin/bash
# tile-shape sweep of the scan kernel through the NVRTC path (SD_TUNE_* change the plan signature)
ROWS=${1:-200000000}
for q in rd q6 q1; do
  for st in 0 1; do
    for rpt in 2 4 8; do
      for mc in 1 2 3; do
        echo -n "$q staged=$st rpt=$rpt ctas=$mc : "
        SD_TUNE_STAGES=$st SD_TUNE_RPT=$rpt SD_TUNE_MIN_CTAS=$mc timeout 120 python tools/profile_scan.py $q $ROWS 4 2>&1 | tail -1
      done
    done
  done
done
