#!/bin/bash
cd "$(dirname "$0")/.."
echo "== shipped, one block of 512 (shade_wide=1)"; SHADE_WIDE=1 timeout 300 python tools/shade_cap_ab.py 2>&1 | tail -5
for l in wb768 wb1024; do
  echo "== $l: one block of that many threads (shade_wide=0 = the library's WF_BLOCK)"
  SHADE_WIDE=0 B200_CYCLES_LIB=$PWD/raytracingproject_b200/_build/lib_$l.so timeout 300 python tools/shade_cap_ab.py 2>&1 | tail -5
done
