#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
B="--no-cpu-baseline --no-e2e --configs none"
echo "== cornell principled 128 spp"
BENCH_ARGS="--workload cornell --spp 128" tools/variants.sh run old base
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade_surface --launch-skip 3 --launch-count 2 \
  -o $O/r02c_shade -f python bench.py --workload cornell --spp 32 --steps 1 --warmup 0 $B > $O/ncu_c.log 2>&1
tail -3 $O/ncu_c.log | cut -c1-300
