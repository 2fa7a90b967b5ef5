#!/bin/bash
# round 2, GPU call P (8 GPUs): the NCCL exchange test on 2 of them, then the strong-scaling bench on all 8
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -s 2>&1 | tail -25 > gpurun_out/p_multirank.txt
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 --no-extras --no-cpu > gpurun_out/p_bench_n$N.json 2> gpurun_out/p_bench_n$N.err
done
cat gpurun_out/p_multirank.txt
python - <<'PY'
import json
for N in (8,4):
    try:
        d=json.loads(open('gpurun_out/p_bench_n%d.json'%N).read().strip().splitlines()[-1])
        print(N, 'value', d['value'], 'ms', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms_per_launch'], 'frac', d['roofline']['frac'], 'parity', d.get('parity_check',{}).get('ok'), d.get('parity_check',{}).get('max_rel_err'), 'exchange', d.get('exchange'), 'also', d.get('also',{}).get('value'))
    except Exception as e: print(N, 'ERR', e); print(open('gpurun_out/p_bench_n%d.err'%N).read()[-1500:])
PY
