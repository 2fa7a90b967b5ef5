#!/bin/bash
# round 2, GPU call O: the ring with the generic->async proxy fence before the release (default now)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/o_pytest.txt
( echo "== shipped kernel (fence.proxy.async before the release), 60 M rows"; python tools/hash_diag.py 300 3 2>&1 | grep -E " run |verify"
  echo "== shipped kernel, verify build, 6 M rows"; SD_DEBUG_VERIFY=1 SD_JIT_DEFINES="-DSD_EXP_VERIFY=1" python tools/hash_diag.py 30 2 2>&1 | grep -E " run |verify\] mism|verify\] of"
  echo "== WITHOUT the fence (-DSD_EXP_NO_PROXY_FENCE), 6 M rows"; SD_JIT_DEFINES="-DSD_EXP_NO_PROXY_FENCE" python tools/hash_diag.py 30 2 2>&1 | grep -E " run "
  echo "== WITHOUT the fence, verify build, 6 M rows"; SD_DEBUG_VERIFY=1 SD_JIT_DEFINES="-DSD_EXP_VERIFY=1 -DSD_EXP_NO_PROXY_FENCE" python tools/hash_diag.py 30 2 2>&1 | grep -E " run |verify\] mism|verify\] of"
) > gpurun_out/o_hash_diag.txt 2>&1
timeout 500 compute-sanitizer --tool racecheck python tools/hash_diag.py 2 1 > gpurun_out/o_racecheck_hash.txt 2>&1
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err
tail -6 gpurun_out/o_pytest.txt; cat gpurun_out/o_hash_diag.txt; grep -E " run |RACECHECK SUMMARY|hazards\]" gpurun_out/o_racecheck_hash.txt | head -8
python - <<'PY'
import json
d=json.loads(open('gpurun_out/o_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'parity', d['parity_check']['ok'], 'also', d.get('also',{}).get('value'))
PY
