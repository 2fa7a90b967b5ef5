#!/bin/bash
# round 2, GPU call W (2 GPUs): the exchange test and a lean N = 2 bench with parity_check, on the final tree
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multirank.py -m gpu -q -s 2>&1 | tail -22 > gpurun_out/w_multirank.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e --no-cpu --no-extras > gpurun_out/w_bench_n2.json 2> gpurun_out/w_bench_n2.err
cat gpurun_out/w_multirank.txt | tail -14
python - <<'PY'
import json
d=json.loads(open('gpurun_out/w_bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms_per_launch'], 'parity', d['parity_check']['ok'], d['parity_check'].get('max_rel_err'), 'also', d['also']['value'], d['also'].get('parity_check',{}).get('ok'), d.get('exchange'))
PY
