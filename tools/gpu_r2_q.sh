#!/bin/bash
# round 2, GPU call Q (8 GPUs, lean): where does the per-query time go at N = 8?  (host-side laps, SD_DEBUG_TIMING)
mkdir -p gpurun_out
SD_DEBUG_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 128 --warmup 8 --no-e2e --no-cpu --no-also --no-parity --no-extras > gpurun_out/q_bench_n8.json 2> gpurun_out/q_bench_n8.err
grep -E "rank 0\]|rank 7\]|rank 3\]" gpurun_out/q_bench_n8.err | tail -12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q_bench_n8.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms_per_launch'])
PY
