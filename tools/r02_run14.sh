#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for opt in "loose_instances=1" "instance_detail_boxes=16" "instance_detail_boxes=64" "instance_detail_boxes=256" "instance_detail_boxes=1024" "instance_detail_boxes=4096"; do
python bench.py --steps 3 --warmup 3 --workload instanced --spp 16 --no-cpu-baseline --no-e2e --configs none --opt $opt 2>>$O/r02j.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$opt', 'Mrays/s %.0f ms %.1f frac %.3f nodes/ray %.2f tris/ray %.2f inst/ray %.2f closest Grays/s %.3f shadow inst/ray %.2f build_ms %.0f' % (d['value'], d['ms_per_step'], r['frac'], r['nodes_per_ray'], r['tris_per_ray'], r['instances_per_ray'], r['grays_per_s'], r['shadow']['instances_per_ray'], d['config']['bvh8']['build_ms']))"
done
