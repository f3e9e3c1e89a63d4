#!/bin/bash
# L2 persisting access-policy window over the BVH8 node array: A/B
cd "$(dirname "$0")/.."
for wl in "--spp 64" "--workload instanced --spp 16" "--workload cornell --spp 128"; do
echo "== $wl"
for v in 0 1 0 1; do
python bench.py --steps 3 --warmup 3 $wl --no-cpu-baseline --no-e2e --configs none --opt l2_persist_nodes=$v 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('l2_persist_nodes=$v', 'Mrays/s %.0f ms %.1f frac %.3f closest Grays/s %.3f shadow GB/s %.0f node MB %.1f' % (d['value'], d['ms_per_step'], r['frac'], r['grays_per_s'], r['shadow']['achieved'], d['config']['bvh8']['node_bytes']/1e6))"
done
done
