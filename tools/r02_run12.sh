#!/bin/bash
# tile-local sort by shader: GPU tests, then A/B against the global counting sort
cd "$(dirname "$0")/.."
O=gpurun_out
( time python -m pytest tests -m gpu -q -x ) > $O/r02i_pytest.log 2>&1
tail -4 $O/r02i_pytest.log | cut -c1-300
show() {
python -c "
import json,sys
d=json.loads(sys.stdin.read())
c=d['control']
print('$1', 'Mrays/s %.0f ms/step %.2f' % (d['value'], d['ms_per_step']), 'launches %d' % c['launches_per_step'], 'share', {k: round(v,3) for k,v in c['share'].items()})
"
}
for wl in "--workload cube" "--workload cornell --spp 128" "--workload terrain --spp 64" "--workload instanced --spp 16"; do
  echo "== $wl"
  for g in 1 0 1 0; do
    python bench.py --steps 3 --warmup 3 $wl --no-cpu-baseline --no-e2e --configs none --opt sort_global=$g 2>>$O/r02i_ab.err | show "sort_global=$g"
  done
done
