#!/bin/bash
# usage: tools_tune.sh "<opt list variants separated by ;>"
for v in "$@"; do
  opts=""
  for o in $v; do opts="$opts --opt $o"; done
  python bench.py --steps 2 --warmup 1 --spp 64 --no-cpu-baseline --no-e2e $opts 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('$v', 'Mrays/s %.0f' % d['value'], 'closest Grays/s %.3f frac %.3f share %.2f' % (r['grays_per_s'], r['frac'], r['share_of_step']), 'shadow share %.2f' % r['shadow']['share_of_step'], 'ms/step %.1f' % d['ms_per_step'])
"
done
