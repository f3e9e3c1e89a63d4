#!/bin/bash
cd "$(dirname "$0")/.."
for l in base wb512; do
  lib=$PWD/raytracingproject_b200/_build/lib_$l.so
  [ $l = base ] && lib=$PWD/raytracingproject_b200/libb200cycles.so
  echo "== $l full kernels (shade_dense=0)"; SHADE_DENSE=0 B200_CYCLES_LIB=$lib timeout 300 python tools/shade_cap_ext_ab.py 2>&1 | tail -4
  echo "== $l lean GGX kernel"; DIST=GGX SHADE_DENSE=0 B200_CYCLES_LIB=$lib timeout 300 python tools/shade_cap_ab.py 2>&1 | tail -5
done
echo "== instanced 16 spp"
BENCH_ARGS="--workload instanced --spp 16" tools/variants.sh run base wb512
