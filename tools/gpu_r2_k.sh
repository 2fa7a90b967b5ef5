#!/bin/bash
# round 2, GPU call K: what makes the global-table MODE_HASH path mis-attribute rows at scale (call J: every group off, totals kept)?
mkdir -p gpurun_out
export SD_TUNE_NO_FRONT_TABLE=1
( echo "== baseline";            python tools/hash_diag.py 30 2 | grep run
  echo "== no ring (STAGES=0)";  SD_TUNE_STAGES=0 python tools/hash_diag.py 30 2 | grep run
  echo "== 1 CTA/SM";            SD_TUNE_MIN_CTAS=1 python tools/hash_diag.py 30 2 | grep run
  echo "== RPT=2";               SD_TUNE_RPT=2 python tools/hash_diag.py 30 2 | grep run
  echo "== RPT=2, 1 CTA/SM";     SD_TUNE_RPT=2 SD_TUNE_MIN_CTAS=1 python tools/hash_diag.py 30 2 | grep run
  echo "== 3 batches";           python tools/hash_diag.py 3 2 | grep run
) > gpurun_out/k_hash_diag.txt 2>&1
timeout 500 compute-sanitizer --tool memcheck python tools/hash_diag.py 2 1 > gpurun_out/k_memcheck.txt 2>&1
timeout 500 compute-sanitizer --tool initcheck python tools/hash_diag.py 2 1 > gpurun_out/k_initcheck.txt 2>&1
timeout 500 compute-sanitizer --tool racecheck python tools/hash_diag.py 2 1 > gpurun_out/k_racecheck.txt 2>&1
unset SD_TUNE_NO_FRONT_TABLE
python -m pytest tests/test_gpu_general.py -m gpu -q -k "written_on_the_device or grows_between" 2>&1 | tail -30 > gpurun_out/k_pytest.txt
cat gpurun_out/k_hash_diag.txt
for f in memcheck initcheck racecheck; do echo "-- $f"; grep -v "^=========\s*$" gpurun_out/k_$f.txt | grep -E "run 0|ERROR SUMMARY|RACECHECK SUMMARY|Error|Uninit|hazard|at .*sd_" | head -14; done
tail -25 gpurun_out/k_pytest.txt
