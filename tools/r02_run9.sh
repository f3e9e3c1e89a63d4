#!/bin/bash
# pipelined bounce loop: GPU tests, then A/B against a stream sync per iteration
cd "$(dirname "$0")/.."
O=gpurun_out
( time python -m pytest tests -m gpu -q -x ) > $O/r02g_pytest.log 2>&1
tail -5 $O/r02g_pytest.log | cut -c1-300
show() {
python -c "
import json,sys
d=json.loads(sys.stdin.read())
c=d['control']
print('$1', 'Mrays/s %.0f ms/step %.2f' % (d['value'], d['ms_per_step']), 'launches %d iters %.0f syncs %.0f waits %.0f' % (c['launches_per_step'], c['iterations_per_step'], c['stream_syncs_per_step'], c['lagged_counter_reads_per_step']), 'share', {k: round(v,3) for k,v in c['share'].items()})
"
}
for wl in "--workload cube" "--workload cornell --spp 128" "--workload terrain --spp 64" "--workload instanced --spp 16"; do
  echo "== $wl"
  for sync in 1 0 1 0; do
    python bench.py --steps 3 --warmup 3 $wl --no-cpu-baseline --no-e2e --configs none --opt sync_iterations=$sync 2>>$O/r02g_ab.err | show "sync_iterations=$sync"
  done
done
