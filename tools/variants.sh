#!/bin/bash
# Build compile-time variants of libb200cycles.so for A/B runs on the GPU box.
#   tools/variants.sh build name1 "-DFOO=1" name2 "-DBAR=2" ...   (here, no GPU needed)
#   tools/variants.sh run name1 name2 ...                         (on the box; prints one line each)
set -e; set +e
cd "$(dirname "$0")/.."
mode=$1; shift
B=raytracingproject_b200/_build
mkdir -p $B
if [ "$mode" = build ]; then
  while [ $# -gt 0 ]; do
    name=$1; flags=$2; shift 2
    make -s -C raytracingproject_b200/csrc OUT=../_build/lib_$name.so EXTRA="$flags" 2>&1 | grep -E "error|spill stores" | grep -v " 0 bytes spill stores" | head -5 || true
    echo "built $name ($flags)"
  done
else
  for name in "$@"; do
    lib=$PWD/$B/lib_$name.so
    [ "$name" = base ] && lib=$PWD/raytracingproject_b200/libb200cycles.so
    B200_CYCLES_LIB=$lib python bench.py --steps 3 --warmup 2 ${BENCH_ARGS:---spp 64} --no-cpu-baseline --no-e2e --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('$name', 'Mrays/s %.0f' % d['value'], 'closest Grays/s %.3f frac %.3f share %.2f' % (r['grays_per_s'], r['frac'], r['share_of_step']), 'shadow GB/s %.0f share %.2f' % (r['shadow']['achieved'], r['shadow']['share_of_step']), 'ms/step %.1f' % d['ms_per_step'])
"
  done
fi
