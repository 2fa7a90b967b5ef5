#!/bin/bash
# round 2, GPU call B: LZ4 throughput at high residency + e2e_lz4 with the variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/b_pytest.txt
( LZ4_KINDS=0,1 LZ4_VARIANTS=2,3 python tools/lz4_bench.py 3000 200000 3; LZ4_KINDS=0 LZ4_VARIANTS=2,3 python tools/lz4_bench.py 6000 200000 3 ) > gpurun_out/b_lz4.txt 2>&1
for cfg in "0 0 256" "1 1 256" "1 1 1024" "0 1 1024" "1 1 4096"; do
  set -- $cfg
  echo "== dense=$1 parse=$2 flush_mb=$3" >> gpurun_out/b_e2e.txt
  SD_TUNE_LZ4_DENSE=$1 SD_TUNE_LZ4_PARSE=$2 SD_TUNE_FLUSH_MB=$3 BENCH_DEBUG=1 python bench.py --rows 200000000 --no-cpu --no-also --no-parity --steps 5 2>> gpurun_out/b_e2e.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:(d[k]['value']/1e9, d[k]['ms_per_step']) for k in ('e2e','e2e_lz4')})" >> gpurun_out/b_e2e.txt
done
tail -25 gpurun_out/b_pytest.txt; cat gpurun_out/b_lz4.txt; grep -v "^\[rank" gpurun_out/b_e2e.txt | grep -v "e2e step" ; grep "e2e step" gpurun_out/b_e2e.txt | tail -4
