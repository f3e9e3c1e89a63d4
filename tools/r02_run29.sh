#!/bin/bash
cd "$(dirname "$0")/.."
for w in instanced terrain; do
for t in 0 12 16 20 28 31; do
  echo -n "$w refill_threshold=$t: "
  python bench.py --steps 2 --warmup 1 --workload $w --spp 16 --no-cpu-baseline --no-e2e --configs none --opt refill_threshold=$t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('Mrays/s %.0f closest Grays/s %.3f frac %.3f shadow GB/s %.0f' % (d['value'], r['grays_per_s'], r['frac'], r['shadow']['achieved']))
"
done
done
