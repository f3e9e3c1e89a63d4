#!/bin/bash
# two real GPUs: the one-process MultiDevice path (NCCL film all-reduce inside the C ABI),
# then the driver-style N=2 bench (terrain strong split + config 5)
cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi -L
( time python -m pytest tests/test_device_shim_gpu.py -m gpu -q -s -k "multi_device or reference_scene" ) > $O/r02g_multidevice_2gpu.log 2>&1
tail -6 $O/r02g_multidevice_2gpu.log | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02g_bench_n2.json 2> $O/r02g_bench_n2.err
tail -c 300 $O/r02g_bench_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02g_bench_n2.json').read().strip().splitlines()[-1])
print('N=2', d['value'], d['ms_per_step'], d.get('allreduce_ms'), {k:(v.get('value'), v.get('ms_per_step'), v.get('allreduce_ms')) for k,v in (d.get('configs') or {}).items()})
P
