#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
run() { # name lib opts...
  lib=$PWD/raytracingproject_b200/libb200cycles.so
  [ "$2" != base ] && lib=$PWD/raytracingproject_b200/_build/lib_$2.so
  B200_CYCLES_LIB=$lib python bench.py --steps 3 --warmup 3 $BENCH_ARGS --no-cpu-baseline --no-e2e --configs none $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['control']['share']
print('$1', 'Mrays/s %.0f ms/step %.1f' % (d['value'], d['ms_per_step']), {k: round(v,3) for k,v in c.items()})"
}
for wl in "--workload cornell --spp 128" "--workload cube" "--spp 64"; do
  echo "== $wl"; export BENCH_ARGS="$wl"
  run base base ""
  run wb128 wb128 ""
  run carve25 base "--opt shade_carveout=25"
  run carve50 base "--opt shade_carveout=50"
  run carve100 base "--opt shade_carveout=100"
  run base base ""
done
