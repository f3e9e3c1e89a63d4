#!/bin/bash
cd "$(dirname "$0")/.."
for l in base mb3b; do
  lib=$PWD/raytracingproject_b200/_build/lib_$l.so
  [ $l = base ] && lib=$PWD/raytracingproject_b200/libb200cycles.so
  echo "== $l"; B200_CYCLES_LIB=$lib timeout 300 python tools/shade_cap_ext_ab.py 2>&1 | tail -5
done
