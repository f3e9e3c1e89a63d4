#!/bin/bash
# driver-style scaling point at N GPUs of one box: reference arm (rank 0 only) and our arm
cd "$(dirname "$0")/.."
O=gpurun_out; N=${1:-4}
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > $O/r02p_bench_reference_n$N.json 2> $O/r02p_ref_n$N.err
tail -c 200 $O/r02p_ref_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 3 --warmup 3 > $O/r02p_bench_n$N.json 2> $O/r02p_bench_n$N.err
tail -c 300 $O/r02p_bench_n$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02p_bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N value', d['value'], 'ms', d['ms_per_step'], 'allreduce_ms', d.get('allreduce_ms'), 'e2e', (d.get('e2e') or {}).get('value'))
print('config5', d.get('config5'))
r=json.loads(open('gpurun_out/r02p_bench_reference_n$N.json').read().strip().splitlines()[-1])
print('reference', r.get('value'), r.get('cpu_baseline',{}).get('cores'))
P
