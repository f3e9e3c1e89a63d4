#!/bin/bash
# BVH8 host layout + image textures + whole GPU suite, then the driver-style bench line
cd "$(dirname "$0")/.."
O=gpurun_out
( time python -m pytest tests -m gpu -q -x ) > $O/r02h_pytest.log 2>&1
tail -5 $O/r02h_pytest.log | cut -c1-300
python -m pytest tests/test_device_shim_gpu.py -m gpu -q -s -k reference_scene 2>&1 | grep -E "bvh|passed|failed" | cut -c1-400
python bench.py --steps 3 --warmup 3 > $O/r02h_bench.json 2> $O/r02h_bench.err
tail -c 300 $O/r02h_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02h_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['reference_flow'])
print('control', d['control'])
for k,v in (d.get('configs') or {}).items():
    print(k, round(v['value']), round(v['ms_per_step'],1), 'frac', round(v.get('roofline',{}).get('frac',0),3), v.get('control',{}).get('share'))
print('config5', d.get('config5'))
print('cpu', d['cpu_baseline']['value'])
P
