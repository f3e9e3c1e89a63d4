#!/bin/bash
cd "$(dirname "$0")/.."
echo "== mb3 lib"; B200_CYCLES_LIB=$PWD/raytracingproject_b200/_build/lib_mb3.so timeout 300 python tools/shade_cap_ab.py 2>&1 | tail -5
for m in 0 1 -1; do
  echo "== shipped lib shade_dense=$m"; SHADE_DENSE=$m timeout 300 python tools/shade_cap_ab.py 2>&1 | tail -5
done
