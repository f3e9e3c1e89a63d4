#!/bin/bash
# tighter instance bounds in the TLAS: hit-id parity on instanced scenes, then config 4
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_traversal_gpu.py tests/test_golden_gpu.py tests/test_full_gates_gpu.py tests/test_device_shim_gpu.py -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --workload instanced --spp 16 --no-cpu-baseline --no-e2e --configs none > $O/r02j_instanced.json 2>$O/r02j.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02j_instanced.json').read().strip().splitlines()[-1])
r=d['roofline']
print('instanced Mrays/s %.0f ms %.1f frac %.3f nodes/ray %.2f tris/ray %.2f inst/ray %.2f bytes/ray %.0f grays %.3f bvh %s' % (d['value'], d['ms_per_step'], r['frac'], r['nodes_per_ray'], r['tris_per_ray'], r['instances_per_ray'], r['bytes_per_ray'], r['grays_per_s'], d['config']['bvh8']))
print('shadow', r['shadow'])
print(d['control']['share'])
P
